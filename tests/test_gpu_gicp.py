"""GPU parity tests (run on a B200 through gpurun): the CUDA path, called through the C-ABI,
against the CPU oracle on the same seeded inputs.  Bars (BASELINE.json north_star):
kNN indices bit-exact; H/b within 1e-4 relative (we assert 1e-9); poses within 1e-4 m / 1e-5 rad."""
import numpy as np
import pytest

from conftest import rot_angle

pytestmark = pytest.mark.gpu

FLT_MAX = float(np.finfo(np.float32).max)


@pytest.fixture(scope="module")
def rgc():
    import rgc_slam_b200
    return rgc_slam_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def test_extension_is_loaded(rgc):
    import ctypes
    L = rgc.lib()
    assert isinstance(L, ctypes.CDLL)
    ctx = rgc.Context(0)
    assert ctx.launch_count == 0
    ctx.close()


@pytest.mark.parametrize("k", [1, 5, 20, 32])
def test_knn_bitexact_self(rgc, orc, scan_pair, k):
    src, tgt, _ = scan_pair
    idx, d2 = rgc.knn(tgt, tgt, k)
    oi, od = orc.knn(tgt, tgt, k)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2, od)


@pytest.fixture
def knn_defer(rgc):
    """sets the candidate count at which the tile kernel hands a tile to the warp-per-query kernel"""
    import ctypes as C
    from rgc_slam_b200 import api
    L = api.lib()
    L.rgc_debug_set_knn_defer.argtypes = [C.c_void_p, C.c_int]
    ctx = api.default_context()

    def set_(v):
        ctx.check(L.rgc_debug_set_knn_defer(ctx._h, v))
    yield set_
    set_(-1)  # back to the size-dependent default


@pytest.mark.parametrize("defer", [600, 1, 40, 0, -1])
@pytest.mark.parametrize("k", [1, 7, 20, 32])
def test_knn_self_tile_kernel_bitexact(rgc, orc, scan_pair, k, defer, knn_defer):
    """the production self-kNN (warp-cooperative tile kernel used by calculate_covariances);
    defer=1 sends nearly every tile through the warp-per-query kernel, 0 none; -1 is the size-dependent default
    (sweep-sized clouds: the warp-per-query kernel for every point, no tile kernel at all)"""
    knn_defer(defer)
    src, tgt, _ = scan_pair
    for cloud in (tgt, src[:1000], src[:33], src[:5]):
        idx = rgc.knn_self(cloud, k)
        oi, _ = orc.knn(cloud, cloud, k)
        assert np.array_equal(idx, oi)
    rng = np.random.default_rng(3)
    g = np.stack(np.meshgrid(*[np.arange(11)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    L = np.ones((len(g), 4), np.float32)
    L[:, :3] = g[rng.permutation(len(g))]          # integer lattice: every distance tied many times
    oi, _ = orc.knn(L, L, k, brute=True)
    assert np.array_equal(rgc.knn_self(L, k), oi)
    C = np.ones((4000, 4), np.float32)
    C[:, :3] = np.concatenate([rng.normal(0, 0.05, (2000, 3)), rng.normal(0, 30, (2000, 3))]).astype(np.float32)  # dense blob + sparse halo
    oi, _ = orc.knn(C, C, k, brute=True)
    assert np.array_equal(rgc.knn_self(C, k), oi)


def test_knn_bitexact_cross_and_far(rgc, orc, scan_pair):
    src, tgt, _ = scan_pair
    idx, d2 = rgc.knn(tgt, src, 1)
    oi, od = orc.knn(tgt, src, 1)
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)
    rng = np.random.default_rng(0)
    P = np.ones((5000, 4), np.float32)
    P[:, :3] = rng.normal(0, 10, (5000, 3))
    Q = np.ones((3000, 4), np.float32)
    Q[:, :3] = rng.normal(0, 60, (3000, 3))  # most queries far outside the grid
    for k in (1, 20):
        idx, d2 = rgc.knn(P, Q, k)
        oi, od = orc.knn(P, Q, k, brute=True)
        assert np.array_equal(idx, oi) and np.array_equal(d2, od)


def test_knn_edge_cases(rgc, orc):
    rng = np.random.default_rng(1)
    # fewer points than k, with exact duplicates (ties broken by index)
    P = np.ones((7, 4), np.float32)
    P[:, :3] = rng.normal(0, 1, (7, 3))
    P[3] = P[2]
    P[5] = P[2]
    idx, d2 = rgc.knn(P, P, 20)
    oi, od = orc.knn(P, P, 20, brute=True)
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)
    assert (idx[:, 7:] == -1).all() and np.isinf(d2[:, 7:]).all()
    # single point
    idx, d2 = rgc.knn(P[:1], P, 1)
    assert (idx == 0).all()
    # integer lattice: massive exact distance ties, order must be (d2, index)
    g = np.stack(np.meshgrid(*[np.arange(12)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    L = np.ones((len(g), 4), np.float32)
    L[:, :3] = g[rng.permutation(len(g))]
    idx, d2 = rgc.knn(L, L, 20)
    oi, od = orc.knn(L, L, 20, brute=True)
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)
    # point strides 16 / 32 / 48 bytes (PointXYZ / PointXYZI / PointNormal)
    for floats in (4, 8, 12):
        W = np.zeros((len(L), floats), np.float32)
        W[:, :3] = L[:, :3]
        W[:, 3:] = 7.0
        i2, _ = rgc.knn(W, W, 5)
        assert np.array_equal(i2, oi[:, :5])


@pytest.mark.parametrize("method", [0, 1, 2, 3, 4])
def test_covariances(rgc, orc, small_pair, method):
    src, tgt, _ = small_pair
    g = rgc.FastGICP()
    g.setRegularizationMethod(method)
    g.setInputSource(src)
    g.setInputTarget(tgt)
    c = g.getTargetCovariances()
    oi, _ = orc.knn(tgt, tgt, 20)
    oc = orc.covariances_from_knn(tgt, oi, method)
    scale = np.abs(oc).max(axis=(1, 2), keepdims=True)
    assert (np.abs(c - oc) / scale).max() < 1e-8
    assert (c[:, 3, :] == 0).all() and (c[:, :, 3] == 0).all()


@pytest.mark.parametrize("thr", [FLT_MAX, 2.0, 0.3])
def test_linearize_and_compute_error(rgc, orc, scan_pair, thr):
    src, tgt, _ = scan_pair
    g = rgc.FastGICP()
    g.setMaxCorrespondenceDistance(thr)
    g.setInputTarget(tgt)
    g.setInputSource(src)
    o = orc.FastGICP(corr_dist=thr)
    o.setInputTarget(tgt)
    o.setInputSource(src)
    T = np.eye(4)
    T[:3, 3] = [0.1, -0.05, 0.02]
    c, s = np.cos(0.01), np.sin(0.01)
    T[:3, :3] = [[c, -s, 0], [s, c, 0], [0, 0, 1]]
    e, H, b = g.linearize(T)
    oe, oH, ob = o.linearize(T)
    corr, d2 = g.correspondences()
    ocorr, od2 = o.correspondences()
    assert np.array_equal(corr, ocorr)
    inl = ocorr >= 0
    assert np.array_equal(d2[inl], od2[inl])
    assert abs(e - oe) <= 1e-9 * abs(oe)
    assert np.abs(H - oH).max() <= 1e-9 * np.abs(oH).max()
    assert np.abs(b - ob).max() <= 1e-9 * np.abs(ob).max()
    # cost only (H = b = nullptr)
    assert abs(g.linearize(T, want_Hb=False) - oe) <= 1e-9 * abs(oe)
    # compute_error at a trial pose reuses the frozen correspondences / Mahalanobis
    T2 = T.copy()
    T2[:3, 3] += [0.02, 0.01, -0.01]
    assert abs(g.compute_error(T2) - o.compute_error(T2)) <= 1e-9 * abs(oe)
    # bit-reproducible run to run (fixed-order reduction)
    e2, H2, b2 = g.linearize(T)
    assert e2 == e and np.array_equal(H2, H) and np.array_equal(b2, b)


def _pose_close(T, To):
    assert np.abs(T[:3, 3].astype(np.float64) - To[:3, 3]).max() < 1e-4
    assert rot_angle(T[:3, :3], To[:3, :3]) < 1e-5


def test_align_c1_defaults(rgc, orc, scan_pair):
    """Config C1: scan-to-scan, library defaults (k=20, PLANE, LM, 64 it, corr inf)."""
    src, tgt, Ttrue = scan_pair
    g = rgc.FastGICP()
    g.setInputTarget(tgt)
    g.setInputSource(src)
    T = g.align(want_output=True)
    o = orc.FastGICP()
    o.setInputTarget(tgt)
    o.setInputSource(src)
    To = o.align(want_points=True)
    _pose_close(T, To)
    assert g.hasConverged() == o.last["converged"]
    assert g.last_result["iterations"] == o.last["iterations"]
    assert g.last_result["n_linearize"] == o.last["n_linearize"]
    assert g.last_result["n_compute_error"] == o.last["n_compute_error"]
    assert np.abs(g.getFinalHessian() - o.last["final_hessian"]).max() <= 1e-8 * np.abs(o.last["final_hessian"]).max()
    assert np.abs(g.output - o.last["points"]).max() < 2e-4
    assert np.abs(T[:3, 3] - Ttrue[:3, 3]).max() < 0.05  # and it is actually the right answer
    fs, ofs = g.getFitnessScore(), o.getFitnessScore()
    assert abs(fs - ofs) <= 1e-6 * ofs
    assert abs(g.getFitnessScore(0.05) - o.getFitnessScore(0.05)) <= 1e-6 * ofs


def test_align_call_site_params_and_guess(rgc, orc, small_pair):
    """Parameters of the odometer call site (RGC_odometer.cpp:1000-1006) with a non-identity guess."""
    src, tgt, Ttrue = small_pair
    guess = np.eye(4, dtype=np.float32)
    guess[:3, 3] = [0.2, 0.05, 0.0]
    for opt in (1, 0):
        g = rgc.FastGICP()
        g.setMaximumIterations(25)
        g.setMaxCorrespondenceDistance(2.0)
        g.setTransformationEpsilon(1e-6)
        g.setOptimizer(opt)
        g.setInputTarget(tgt)
        g.setInputSource(src)
        T = g.align(guess)
        o = orc.FastGICP(max_iterations=25, corr_dist=2.0, transformation_epsilon=1e-6, optimizer=opt)
        o.setInputTarget(tgt)
        o.setInputSource(src)
        To = o.align(guess)
        _pose_close(T, To)
        assert g.last_result["iterations"] == o.last["iterations"]


def test_identity_caching_swap_and_user_covariances(rgc, orc, small_pair):
    src, tgt, _ = small_pair
    g = rgc.FastGICP()
    g.setInputTarget(tgt)
    g.setInputSource(src)
    T1 = g.align()
    n0 = g.ctx.launch_count
    g.setInputTarget(tgt)  # same object -> early return, nothing rebuilt (fast_gicp_impl.hpp:84-86)
    g.setInputSource(src)
    assert g.ctx.launch_count == n0
    T2 = g.align()
    assert np.array_equal(T1, T2)
    # swapSourceAndTarget: aligning the other way gives (approximately) the inverse
    g.swapSourceAndTarget()
    Tinv = g.align()
    o = orc.FastGICP()
    o.setInputTarget(src)
    o.setInputSource(tgt)
    _pose_close(Tinv, o.align())
    # user-supplied covariances replace the estimated ones
    g2 = rgc.FastGICP()
    g2.setInputTarget(tgt)
    g2.setInputSource(src)
    eye = np.zeros((len(src), 4, 4))
    eye[:, :3, :3] = np.eye(3) * 0.5
    g2.setSourceCovariances(eye)
    assert np.array_equal(g2.getSourceCovariances(), eye)
    o2 = orc.FastGICP()
    o2.setInputTarget(tgt)
    o2.setInputSource(src)
    o2.setSourceCovariances(eye)
    _pose_close(g2.align(), o2.align())


def test_errors_are_loud(rgc):
    g = rgc.FastGICP()
    with pytest.raises(rgc.RgcError):
        g.align()  # no clouds
    with pytest.raises(rgc.RgcError):
        g.setCorrespondenceRandomness(129)  # k > 128 unsupported
    with pytest.raises(rgc.RgcError):
        g.setInputSource(np.zeros((0, 4), np.float32))


def test_fake_sharded_target_matches_unsharded(rgc, scan_pair):
    """Config C5 logic on ONE GPU: two slabs of the target held by two registration objects, every
    source point handled by exactly one of them, partial (err, H, b) summed on the host ==
    the unsharded linearize (SURVEY §4 'fake shard' mode; fp64 summation order is the only difference)."""
    import ctypes as C
    from rgc_slam_b200 import sharded
    src, tgt, _ = scan_pair
    corr, cov_halo = 2.0, 4.0
    T = np.eye(4)
    T[:3, 3] = [0.1, -0.05, 0.02]
    full = rgc.FastGICP()
    full.setMaxCorrespondenceDistance(corr)
    full.setInputTarget(tgt)
    full.setInputSource(src)
    e, H, b = full.linearize(T)
    n_in = (full.correspondences()[0] >= 0).sum()
    L = rgc.lib()
    L.rgc_reg_set_owner_slab.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
    for world in (2, 3):
        edges = sharded.slab_boundaries(tgt[:, 0], world)
        es, Hs, bs, ins = 0.0, np.zeros((6, 6)), np.zeros(6), 0
        for r in range(world):
            g = rgc.FastGICP()
            g.setMaxCorrespondenceDistance(corr)
            keep = sharded.slab_select(tgt, 0, edges[r], edges[r + 1], corr + cov_halo)
            g.setInputTarget(np.ascontiguousarray(tgt[keep]))
            g.setInputSource(src)
            big = float(np.finfo(np.float32).max)
            g.ctx.check(L.rgc_reg_set_owner_slab(g._h, 0, max(float(edges[r]), -big), min(float(edges[r + 1]), big)))
            er, Hr, br = g.linearize(T)
            es, Hs, bs = es + er, Hs + Hr, bs + br
            ins += (g.correspondences()[0] >= 0).sum()
        assert ins == n_in
        assert abs(es - e) <= 1e-11 * abs(e)
        assert np.abs(Hs - H).max() <= 1e-11 * np.abs(H).max() and np.abs(bs - b).max() <= 1e-11 * np.abs(b).max()
