"""CPU: the oracle's kNN / covariance / GICP against independent libraries and analytic truths
(SURVEY.md §8c items (i)-(iii)); frozen golden vectors (item (iv)) in tests/golden/."""
import os

import numpy as np
import pytest
from scipy.spatial import cKDTree

from conftest import rot_angle
from oracle import oracle as orc
from rgc_slam_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_kdtree_equals_bruteforce_and_scipy(small_pair):
    src, tgt, _ = small_pair
    for k in (1, 5, 20):
        i1, d1 = orc.knn(tgt, src, k)
        i2, d2 = orc.knn(tgt, src, k, brute=True)
        assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
    dd, ii = cKDTree(tgt[:, :3].astype(np.float64)).query(src[:, :3].astype(np.float64), 20)
    i1, d1 = orc.knn(tgt, src, 20)
    same_set = np.array([set(a) == set(b) for a, b in zip(ii, i1)])
    ties = int((~same_set).sum())  # float32 vs float64 distance ties near the k-th neighbour
    assert ties <= len(src) // 1000, f"{ties} neighbour-set mismatches"
    assert np.allclose(np.sqrt(d1[same_set]), dd[same_set], rtol=1e-5, atol=1e-6)


def test_knn_tie_policy_is_distance_then_index():
    g = np.stack(np.meshgrid(*[np.arange(6)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    P = np.ones((len(g), 4), np.float32)
    P[:, :3] = g
    idx, d2 = orc.knn(P, P, 7)
    for r in range(len(P)):
        key = list(zip(d2[r].tolist(), idx[r].tolist()))
        assert key == sorted(key)
    assert (idx[:, 0] == np.arange(len(P))).all() and (d2[:, 0] == 0).all()


@pytest.mark.parametrize("method", [orc.REG_NONE, orc.REG_MIN_EIG, orc.REG_NORMALIZED_MIN_EIG, orc.REG_PLANE, orc.REG_FROBENIUS])
def test_covariances_against_numpy(small_pair, method):
    src, tgt, _ = small_pair
    idx, _ = orc.knn(tgt, tgt, 20)
    covs = orc.covariances_from_knn(tgt, idx, method)
    rng = np.random.default_rng(0)
    for i in rng.choice(len(tgt), 200, replace=False):
        nb = tgt[idx[i], :3].astype(np.float64)
        C = np.cov(nb.T, bias=True)
        if method == orc.REG_NONE:
            ref = C
        elif method == orc.REG_FROBENIUS:
            Ci = np.linalg.inv(C + 1e-3 * np.eye(3))
            ref = np.linalg.inv(Ci / np.linalg.norm(Ci))
        else:
            U, s, Vt = np.linalg.svd(C)
            if method == orc.REG_PLANE:
                v = np.array([1, 1, 1e-3])
            elif method == orc.REG_MIN_EIG:
                v = np.maximum(s, 1e-3)
            else:
                v = np.maximum(s / s.max(), 1e-3)
            ref = U @ np.diag(v) @ Vt
        assert np.allclose(covs[i][:3, :3], ref, rtol=1e-6, atol=1e-9)
        assert (covs[i][3] == 0).all() and (covs[i][:, 3] == 0).all()


def test_linearize_is_consistent_with_finite_differences(small_pair):
    """b = J^T M e is (half) the gradient of the cost w.r.t. a left-multiplied se(3) step, with the
    correspondences and Mahalanobis matrices frozen (that is what compute_error evaluates)."""
    src, tgt, _ = small_pair
    o = orc.FastGICP()
    o.setInputTarget(tgt)
    o.setInputSource(src)
    T = np.eye(4)
    T[:3, 3] = [0.05, -0.02, 0.01]
    e0, H, b = o.linearize(T)
    assert np.allclose(H, H.T) and np.all(np.linalg.eigvalsh(H) > 0)
    eps = 1e-6
    g = np.zeros(6)
    for j in range(6):
        d = np.zeros(6)
        d[j] = eps
        D = np.eye(4)
        D[:3, :3] = orc.so3_exp(d[:3])
        D[:3, 3] = d[3:]
        Dm = np.eye(4)
        Dm[:3, :3] = orc.so3_exp(-d[:3])
        Dm[:3, 3] = -d[3:]
        g[j] = (o.compute_error(D @ T) - o.compute_error(Dm @ T)) / (2 * eps)
    assert np.allclose(g, 2 * b, rtol=1e-4, atol=1e-4 * np.abs(b).max())
    assert abs(o.compute_error(T) - e0) <= 1e-12 * e0


def test_align_recovers_known_transform_on_noise_free_scene():
    """Analytic KAT: source = target moved by a known rigid transform (same surface samples), so
    align() must return that transform."""
    scene = synth.Scene.make(7, n_boxes=30)
    traj = synth.trajectory(4, seed=5)
    tgt = synth.to_xyz1(synth.lidar_scan(scene, traj[1], n_azimuth=450, seed=1, sigma=0.0))
    rng = np.random.default_rng(0)
    Ttrue = synth.small_perturbation(rng, 0.15, 1.5)
    src = tgt.copy()
    src[:, :3] = (tgt[:, :3].astype(np.float64) - Ttrue[:3, 3]) @ Ttrue[:3, :3]  # inverse transform
    src = src.astype(np.float32)
    o = orc.FastGICP(transformation_epsilon=1e-7, rotation_epsilon=1e-7)
    o.setInputTarget(tgt)
    o.setInputSource(src)
    T = o.align()
    assert np.abs(T[:3, 3] - Ttrue[:3, 3]).max() < 2e-5
    assert rot_angle(T[:3, :3], Ttrue[:3, :3]) < 2e-6
    assert o.getFitnessScore() < 1e-9


def test_golden_vectors():
    """Frozen oracle outputs (tests/golden/make_golden.py): protects the oracle itself from drift."""
    z = np.load(os.path.join(GOLD, "gicp_small.npz"))
    src, tgt = z["src"], z["tgt"]
    idx, d2 = orc.knn(tgt, tgt, 20)
    assert np.array_equal(idx, z["knn_idx"]) and np.array_equal(d2, z["knn_d2"])
    o = orc.FastGICP()
    o.setInputTarget(tgt)
    o.setInputSource(src)
    e, H, b = o.linearize(z["T_lin"])
    assert np.allclose(e, z["lin_err"], rtol=1e-10)
    assert np.allclose(H, z["lin_H"], rtol=1e-9, atol=1e-9 * np.abs(z["lin_H"]).max())
    assert np.allclose(b, z["lin_b"], rtol=1e-9, atol=1e-9 * np.abs(z["lin_b"]).max())
    assert np.array_equal(o.correspondences()[0], z["corr"])
    T = o.align()
    assert np.abs(T - z["T_final"]).max() < 1e-6
    assert o.last["iterations"] == int(z["iterations"])


def test_vgicp_golden_vectors():
    """Frozen FastVGICP oracle outputs (tests/golden/make_golden.py: vgicp_small)."""
    z = np.load(os.path.join(GOLD, "gicp_small.npz"))
    v = np.load(os.path.join(GOLD, "vgicp_small.npz"))
    for name, search in (("d1", orc.DIRECT1), ("d7", orc.DIRECT7)):
        o = orc.FastVGICP(resolution=1.0, search_method=search)
        o.setInputTarget(z["tgt"])
        o.setInputSource(z["src"])
        e, H, b = o.linearize(z["T_lin"])
        assert np.allclose(e, v[f"{name}_err"], rtol=1e-10)
        assert np.allclose(H, v[f"{name}_H"], rtol=1e-9, atol=1e-9 * np.abs(v[f"{name}_H"]).max())
        assert np.allclose(b, v[f"{name}_b"], rtol=1e-9, atol=1e-9 * np.abs(v[f"{name}_b"]).max())
        assert o.num_correspondences() == int(v[f"{name}_ncorr"])
        if name == "d1":
            coords, num, mean, cov = o.voxels()
            order = np.lexsort((coords[:, 2], coords[:, 1], coords[:, 0]))
            assert np.array_equal(coords[order], v["vox_coords"]) and np.array_equal(num[order], v["vox_num"])
            assert np.allclose(mean[order], v["vox_mean"], rtol=0, atol=1e-12) and np.allclose(cov[order], v["vox_cov"], rtol=0, atol=1e-12)
        T = o.align()
        assert np.abs(T - v[f"{name}_T"]).max() < 1e-6 and o.last["iterations"] == int(v[f"{name}_iterations"])
