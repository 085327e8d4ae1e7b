"""CPU: the PRODUCT's host/device headers (rgc_grid.cuh, rgc_math.cuh, rgc_lm.hpp, rgc_preprocess.cuh, rgc_mapping.cuh), compiled by g++
in tests/hostsim and run one simulated thread per point, against the oracle.  This is the no-GPU
check of the search / algebra logic; the real parity tests (-m gpu) call the CUDA path."""
import numpy as np
import pytest

import hostsim
from oracle import oracle as orc

FLT_MAX = float(np.finfo(np.float32).max)


@pytest.mark.parametrize("k", [1, 5, 20, 32])
def test_grid_knn_is_exact_on_a_lidar_sweep(small_pair, k):
    src, tgt, _ = small_pair
    idx, d2 = hostsim.knn(tgt, tgt, k)
    oi, od = orc.knn(tgt, tgt, k)
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)
    idx, d2 = hostsim.knn(tgt, src, k)
    oi, od = orc.knn(tgt, src, k)
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)


@pytest.mark.parametrize("cell", [0.0, 0.2, 1.5])
def test_grid_knn_edge_cases(cell):
    rng = np.random.default_rng(0)
    P = np.ones((5000, 4), np.float32)
    P[:, :3] = rng.normal(0, 10, (5000, 3))
    Q = np.ones((2000, 4), np.float32)
    Q[:, :3] = rng.normal(0, 60, (2000, 3))  # mostly far outside the grid
    for k in (1, 20):
        idx, d2 = hostsim.knn(P, Q, k, cell)
        oi, od = orc.knn(P, Q, k, brute=True)
        assert np.array_equal(idx, oi) and np.array_equal(d2, od)
    # n < k with exact duplicates
    T = np.ones((7, 4), np.float32)
    T[:, :3] = rng.normal(0, 1, (7, 3))
    T[3] = T[2]
    T[5] = T[2]
    idx, d2 = hostsim.knn(T, T, 20, cell)
    oi, od = orc.knn(T, T, 20, brute=True)
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)
    # integer lattice: every distance tied many times
    g = np.stack(np.meshgrid(*[np.arange(10)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    L = np.ones((len(g), 4), np.float32)
    L[:, :3] = g[rng.permutation(len(g))]
    idx, d2 = hostsim.knn(L, L, 20, cell)
    oi, od = orc.knn(L, L, 20, brute=True)
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)
    # huge coordinates / tiny extent / single point
    B = (P * np.float32(1.0)).copy()
    B[:, :3] += np.float32(5000.0)
    idx, d2 = hostsim.knn(B, B[:500], 5, cell)
    oi, od = orc.knn(B, B[:500], 5, brute=True)
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)
    one = np.ones((1, 4), np.float32)
    idx, d2 = hostsim.knn(one, Q[:10], 1, cell)
    assert (idx == 0).all()


def test_search_work_is_density_independent(scan_pair):
    """the octree descent must not degrade on a sweep whose density varies by orders of magnitude"""
    src, tgt, _ = scan_pair
    _, _, st = hostsim.knn(tgt, tgt, 20, want_stats=True)
    nodes, lookups, cands = st
    assert cands < 150 and lookups < 60, st


@pytest.mark.parametrize("method", [0, 1, 2, 3, 4])
def test_covariance_math(small_pair, method):
    src, tgt, _ = small_pair
    idx, _ = orc.knn(tgt, tgt, 20)
    oc = orc.covariances_from_knn(tgt, idx, method)
    sc = hostsim.covs(tgt, idx, method)
    scale = np.abs(oc).max(axis=(1, 2), keepdims=True)
    assert (np.abs(sc - oc) / scale).max() < 1e-8


@pytest.mark.parametrize("thr", [FLT_MAX, 2.0, 0.3])
def test_linearize_math(small_pair, thr):
    src, tgt, _ = small_pair
    o = orc.FastGICP(corr_dist=thr)
    o.setInputTarget(tgt)
    o.setInputSource(src)
    T = np.eye(4)
    T[:3, 3] = [0.1, -0.05, 0.02]
    oe, oH, ob = o.linearize(T)
    ocorr, _ = o.correspondences()
    e, H, b, corr = hostsim.linearize(src, tgt, o.getSourceCovariances(), o.getTargetCovariances(), T, thr)
    assert np.array_equal(corr, ocorr)
    assert abs(e - oe) <= 1e-10 * abs(oe)
    assert np.abs(H - oH).max() <= 1e-10 * np.abs(oH).max()
    assert np.abs(b - ob).max() <= 1e-10 * np.abs(ob).max()


def test_host_lm_helpers():
    L = hostsim.lib()
    rng = np.random.default_rng(0)
    for _ in range(100):
        B = rng.normal(size=(6, 9))
        H = np.ascontiguousarray(B @ B.T + 1e-6 * np.eye(6))
        b = rng.normal(size=6)
        x = np.empty(6)
        L.sim_solve_ldlt6(H.reshape(-1), b, x)
        assert np.allclose(x, orc.ldlt6_solve(H, b), rtol=1e-9, atol=1e-12)
        assert np.allclose(x, np.linalg.solve(H, b), rtol=1e-7, atol=1e-10)
        d = rng.normal(size=6) * 10 ** rng.uniform(-7, 0)
        D = np.empty(16)
        L.sim_se3_delta(d, D)
        D = D.reshape(4, 4)
        assert np.allclose(D[:3, :3], orc.so3_exp(d[:3]), atol=1e-15) and np.allclose(D[:3, 3], d[3:]) and D[3, 3] == 1
    D = np.eye(4)
    D[0, 3] = 4e-4
    assert L.sim_is_converged(D.reshape(-1), 2e-3, 5e-4) == 1
    D[0, 3] = 6e-4
    assert L.sim_is_converged(D.reshape(-1), 2e-3, 5e-4) == 0
    D[0, 3] = 0
    D[0, 1] = 3e-3
    assert L.sim_is_converged(D.reshape(-1), 2e-3, 5e-4) == 0


# ---- pre-step and mapping association: the product's host/device arithmetic on the CPU ----
def _stamped(n, seed, spread=8.0):
    rng = np.random.default_rng(seed)
    P = np.zeros((n, 4), np.float32)
    P[:, :3] = rng.normal(0, spread, (n, 3))
    P[:, 3] = rng.integers(0, 16, n) + np.float32(0.1) * rng.uniform(0, 1, n).astype(np.float32)
    return P


def test_hostsim_deskew_is_the_oracles():
    from oracle import oracle as orc
    P = _stamped(5000, 1, 25.0)
    q = np.array([0.9998, 0.004, -0.011, 0.0125])
    q /= np.linalg.norm(q)
    t = np.array([0.31, -0.04, 0.012])
    assert np.array_equal(hostsim.deskew(P, q, t), orc.deskew(P, q, t))          # same libm, same rounding: bit-exact
    assert np.array_equal(hostsim.deskew(P, [1, 0, 0, 0], [0, 0, 0]), P)


@pytest.mark.parametrize("leaf", [0.2, 0.3, 1.5])
def test_hostsim_voxel_grid_is_the_oracles(leaf):
    from oracle import oracle as orc
    P = _stamped(8000, 2, 5.0)
    assert np.array_equal(hostsim.voxel_grid(P, leaf), orc.voxel_grid(P, leaf))
    far = np.array([[0, 0, 0, 0], [3000, 3000, 3000, 1]], np.float32)
    assert np.array_equal(hostsim.voxel_grid(far, 0.001), far)


def test_hostsim_map_association_matches_oracle():
    from oracle import oracle as orc
    from test_oracle_mapping import _pose, _scene
    corner, surf, rng = _scene(31)
    rot, q, t = _pose(rng)
    fe = np.zeros((500, 4), np.float32)
    fe[:, :3] = rot.inv().apply(corner[rng.choice(len(corner), 500), :3] + rng.normal(0, 0.05, (500, 3)) - t)
    fe[250:, :3] += rng.normal(0, 3, (250, 3))
    fp = np.zeros((800, 4), np.float32)
    fp[:, :3] = rot.inv().apply(surf[rng.choice(len(surf), 800), :3] + rng.normal(0, 0.03, (800, 3)) - t)
    fp[400:, :3] += rng.normal(0, 4, (400, 3))
    v, a, b = hostsim.map_assoc(corner, fe, q, t, plane=False)
    ov, oa, ob = orc.assoc_edges(corner, fe, q, t)
    assert (v != ov).sum() <= 1 and v.sum() > 50
    both = v & ov
    same = np.abs(a[both] - oa[both]).max(1) < 1e-9
    swap = np.abs(a[both] - ob[both]).max(1) < 1e-9
    assert (same | swap).all()
    v, nrm, d = hostsim.map_assoc(surf, fp, q, t, plane=True)
    ov, on, od = orc.assoc_planes(surf, fp, q, t)
    assert (v != ov).sum() <= 1 and v.sum() > 100
    both = v & ov
    assert np.abs(nrm[both] - on[both]).max() < 1e-10 and np.abs(d[both] - od[both]).max() < 1e-9


@pytest.mark.parametrize("k", [1, 2, 5, 20, 21, 32])
def test_heap64_selects_the_k_smallest(k):
    """the tile kernel's packed-key heap (HeapK64): any insertion order, duplicates of the d2 half, k > n"""
    rng = np.random.default_rng(k)
    for n in (0, 1, k - 1, k, k + 1, 5 * k + 3, 1000):
        if n < 0:
            continue
        d2 = rng.integers(0, 50, n).astype(np.uint64)              # many equal distances
        keys = (d2 << np.uint64(32)) | rng.permutation(n).astype(np.uint64)
        exp = np.sort(keys)[:k]
        got = hostsim.heap64_topk(keys, k, stride=int(rng.integers(1, 4)))
        assert np.array_equal(got[: len(exp)], exp)
        assert (got[len(exp):] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
