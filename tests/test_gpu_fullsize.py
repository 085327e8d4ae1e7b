"""GPU tests at BASELINE.json's full C2 size (VLP-16 sweep vs 500 000-point submap): the oracle where it
finishes in seconds, and size-independent properties where it does not (sortedness, self at rank 0,
path independence of the two kNN kernels, exact brute-force checks on a random sample of queries,
recovery of a known transform)."""
import ctypes as C

import numpy as np
import pytest

from conftest import rot_angle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rgc():
    import rgc_slam_b200
    return rgc_slam_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="module")
def c2():
    import bench
    return bench.build_workload(0, bench.N_SUBMAP, 1)[0]


def _d2(P, Q):
    """the reference's float distance ((dx*dx + dy*dy) + dz*dz), every op rounded to float"""
    d = (P[:, None, :3] - Q[None, :, :3]).astype(np.float32)
    return (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]


def _set_defer(rgc, v):
    from rgc_slam_b200 import api
    L = api.lib()
    L.rgc_debug_set_knn_defer.argtypes = [C.c_void_p, C.c_int]
    ctx = api.default_context()
    ctx.check(L.rgc_debug_set_knn_defer(ctx._h, v))


def test_knn_500k_properties_and_sampled_bruteforce(rgc, c2):
    tgt = c2["tgt"]
    n, k = len(tgt), 20
    idx = rgc.knn_self(tgt, k)
    assert idx.shape == (n, k) and idx.min() >= 0 and idx.max() < n
    P = tgt[:, :3]
    diff = P[idx] - P[:, None, :]
    d2 = ((diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]) + diff[..., 2] * diff[..., 2]).astype(np.float32)
    assert (np.diff(d2, axis=1) >= 0).all()                       # ascending
    tie = np.diff(d2, axis=1) == 0
    assert (np.diff(idx, axis=1)[tie] > 0).all()                  # ties by ascending index
    assert (d2[:, 0] == 0).all()                                  # the point itself (or an exact duplicate) first
    assert (idx[:, 0] <= np.arange(n)).all()
    assert all(len(set(r)) == k for r in idx[:: n // 2000])       # no repeated neighbour
    # exact check of a random sample of queries against brute force over all 500k points
    rng = np.random.default_rng(5)
    qs = rng.choice(n, 256, replace=False)
    for lo in range(0, len(qs), 32):
        q = qs[lo:lo + 32]
        D = _d2(tgt[q], tgt)                                       # [32, n] float32
        cand = np.argpartition(D, 96, axis=1)[:, :96]              # superset of the 20 best (ties included)
        for r in range(len(q)):
            kth = np.sort(D[r, cand[r]])[k - 1]
            c = np.flatnonzero(D[r] <= kth)                        # every point tied with the k-th is a candidate
            o = c[np.lexsort((c, D[r, c]))][:k]
            assert np.array_equal(o, idx[q[r]])
    # both production kernels (tile only / heavy tiles deferred to warp-per-query) give the same lists
    try:
        _set_defer(rgc, 0)
        assert np.array_equal(rgc.knn_self(tgt, k), idx)
        _set_defer(rgc, 150)
        assert np.array_equal(rgc.knn_self(tgt, k), idx)
    finally:
        _set_defer(rgc, 600)


def test_c2_linearize_matches_oracle_at_full_size(rgc, orc, c2):
    src, tgt, guess = c2["src"], c2["tgt"], c2["guess"].astype(np.float64)
    g = rgc.FastGICP()
    g.setMaxCorrespondenceDistance(2.0)
    g.setInputTarget(tgt)
    g.setInputSource(src)
    o = orc.FastGICP(corr_dist=2.0)
    o.setInputTarget(tgt)
    o.setInputSource(src)
    e, H, b = g.linearize(guess)
    oe, oH, ob = o.linearize(guess)
    corr, d2 = g.correspondences()
    ocorr, od2 = o.correspondences()
    assert np.array_equal(corr, ocorr)
    assert np.array_equal(d2[ocorr >= 0], od2[ocorr >= 0])
    assert abs(e - oe) <= 1e-9 * abs(oe)
    assert np.abs(H - oH).max() <= 1e-9 * np.abs(oH).max()        # bar: 1e-4 relative
    assert np.abs(b - ob).max() <= 1e-9 * np.abs(ob).max()
    assert np.array_equal(H, H.T) and np.linalg.eigvalsh(H).min() > 0


def test_c2_align_matches_oracle_and_truth(rgc, orc, c2):
    src, tgt, guess, truth = c2["src"], c2["tgt"], c2["guess"], c2["truth"]
    import bench
    g = bench.new_reg(rgc, rgc.api.default_context())
    g.setInputTarget(tgt)
    g.setInputSource(src)
    T = g.align(guess)
    o = orc.FastGICP(max_iterations=25, corr_dist=2.0, transformation_epsilon=1e-6)
    o.setInputTarget(tgt)
    o.setInputSource(src)
    To = o.align(guess)
    assert np.abs(T[:3, 3].astype(np.float64) - To[:3, 3]).max() < 1e-4
    assert rot_angle(T[:3, :3], To[:3, :3]) < 1e-5
    assert g.last_result["iterations"] == o.last["iterations"]
    assert np.abs(T[:3, 3] - truth[:3, 3]).max() < 0.02
    assert abs(g.getFitnessScore() - o.getFitnessScore()) <= 1e-6 * o.getFitnessScore()


def test_known_transform_is_recovered(rgc, c2):
    """source = a subset of the target moved by a known rigid motion: the optimum is that motion,
    with zero residual, whatever the cloud size"""
    tgt = c2["tgt"]
    rng = np.random.default_rng(11)
    sub = tgt[np.sort(rng.choice(len(tgt), 60000, replace=False))].astype(np.float64)
    a = np.deg2rad(1.2)
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    t = np.array([0.25, -0.12, 0.03])
    Tk = np.eye(4)
    Tk[:3, :3], Tk[:3, 3] = R, t
    src = np.ones((len(sub), 4), np.float32)
    src[:, :3] = ((sub[:, :3] - t) @ R).astype(np.float32)        # inverse motion applied to the points
    g = rgc.FastGICP()
    g.setMaxCorrespondenceDistance(2.0)
    g.setTransformationEpsilon(1e-6)
    g.setInputTarget(tgt)
    g.setInputSource(src)
    T = g.align()
    assert g.hasConverged()
    assert np.abs(T[:3, 3] - t).max() < 2e-4
    assert rot_angle(T[:3, :3], R) < 2e-5
    assert g.getFitnessScore() < 1e-6
