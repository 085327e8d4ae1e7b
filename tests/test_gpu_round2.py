"""GPU parity tests added in round 2 (VERDICT r1 "what's weak" 2-3): paths and configurations
that no round-1 test reached — rejected LM steps with the look-ahead linearize, the "lm not converged"
exit, on-demand target covariances, config C4 inputs and a down-scaled config C5 against the ORACLE,
non-finite input, the VGICP align -> getFitnessScore -> align sequence.
Bars (BASELINE.json north_star): poses within 1e-4 m / 1e-5 rad, H/b within 1e-4 relative (we assert 1e-9)."""
import os

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
def eighth_pair(scene, traj):
    """225-azimuth sweeps (~2.5k points): small enough that a far guess leaves few inliers and a
    near-singular H — the regime where the oracle's LM rejects steps."""
    from rgc_slam_b200 import synth
    tgt = synth.to_xyz1(synth.lidar_scan(scene, traj[10], n_azimuth=225, seed=5))
    src = synth.to_xyz1(synth.lidar_scan(scene, traj[11], n_azimuth=225, seed=6))
    return src, tgt


def far_guess(trial: int) -> np.ndarray:
    """the trial-th draw of: rotation U(20, 180) deg about a random axis, translation U(+-5 m)^3"""
    from scipy.spatial.transform import Rotation as Rot
    rng = np.random.default_rng(0)
    for _ in range(trial + 1):
        ang = rng.uniform(20, 180)
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        g = np.eye(4, dtype=np.float32)
        g[:3, :3] = Rot.from_rotvec(np.deg2rad(ang) * axis).as_matrix()
        g[:3, 3] = rng.uniform(-5, 5, 3)
    return g


def _same_run(g, o, T, To, hess=True):
    assert g.hasConverged() == o.last["converged"]
    r = g.last_result
    assert (r["iterations"], r["n_linearize"], r["n_compute_error"]) == (o.last["iterations"], o.last["n_linearize"], o.last["n_compute_error"])
    assert np.abs(T[:3, 3].astype(np.float64) - To[:3, 3]).max() < 1e-4
    assert rot_angle(T[:3, :3], To[:3, :3]) < 1e-5
    if hess:
        Ho = o.last["final_hessian"]
        assert np.abs(g.getFinalHessian() - Ho).max() <= 1e-6 * np.abs(Ho).max()


# trials found by scanning the oracle (corr 1 m, 30 iterations): every one of them has rejected steps
@pytest.mark.parametrize("look_ahead", [True, False])
@pytest.mark.parametrize("trial", [3, 19, 29, 0])
def test_rejected_lm_steps_match_oracle(rgc, orc, eighth_pair, trial, look_ahead, monkeypatch):
    """rho < 0 -> lambda *= nu, nu *= 2, retry (lsq_registration_impl.hpp:155-163): the trial linearize
    issued behind compute_error is thrown away and the buffers are NOT swapped (rgc_gicp.cu step_lm)."""
    src, tgt = eighth_pair
    guess = far_guess(trial)
    o = orc.FastGICP(max_iterations=30, corr_dist=1.0)
    o.setInputTarget(tgt)
    o.setInputSource(src)
    To = o.align(guess)
    assert o.last["n_compute_error"] > o.last["n_linearize"], "the oracle must reject at least one step for this case to mean anything"
    if not look_ahead:
        monkeypatch.setenv("RGC_NO_LOOKAHEAD", "1")
    ctx = rgc.Context(0)  # the look-ahead switch is read when a context is created
    try:
        g = rgc.FastGICP(ctx)
        g.setMaximumIterations(30)
        g.setMaxCorrespondenceDistance(1.0)
        g.setInputTarget(tgt)
        g.setInputSource(src)
        T = g.align(guess)
        _same_run(g, o, T, To)
        # a second align on the same object starts from scratch (lambda reset, lsq_registration_impl.hpp:56)
        T2 = g.align(guess)
        assert np.array_equal(T, T2)
        g = None
    finally:
        ctx.close()


@pytest.mark.parametrize("lm_max", [1, 2])
def test_lm_not_converged_exit(rgc, orc, eighth_pair, lm_max, capfd):
    """lm_max_iterations_ inner trials all rejected -> step_lm returns false -> "lm not converged!!", break
    (lsq_registration_impl.hpp:69-72, :171)."""
    src, tgt = eighth_pair
    found = None
    for trial in (14, 29, 3, 19, 0):
        o = orc.FastGICP(max_iterations=30, corr_dist=1.0, lm_max_iterations=lm_max)
        o.setInputTarget(tgt)
        o.setInputSource(src)
        To = o.align(far_guess(trial))
        # the break leaves iterations short of the limit without convergence
        if not o.last["converged"] and o.last["iterations"] < 29:
            found = (trial, o, To)
            break
    assert found, "no case exits through 'lm not converged'"
    trial, o, To = found
    g = rgc.FastGICP()
    g.setMaximumIterations(30)
    g.setMaxCorrespondenceDistance(1.0)
    g.setLMMaxIterations(lm_max)
    g.setInputTarget(tgt)
    g.setInputSource(src)
    T = g.align(far_guess(trial))
    _same_run(g, o, T, To)
    assert "lm not converged" in capfd.readouterr().err


def test_on_demand_target_covariances_are_bit_identical(rgc, scan_pair):
    """The default mode computes a target covariance the first time the point becomes a correspondence.
    Every such covariance, and therefore err / H / b and the whole align, must equal the eager pass bit for bit."""
    src, tgt, _ = scan_pair

    def make(on_demand):
        g = rgc.FastGICP()
        g.setMaxCorrespondenceDistance(2.0)
        g.setTargetCovarianceMode(on_demand)
        g.setInputTarget(tgt)
        g.setInputSource(src)
        return g

    lazy, eager = make(True), make(False)
    T = np.eye(4)
    T[:3, 3] = [0.1, -0.05, 0.02]
    el, Hl, bl = lazy.linearize(T)
    ee, He, be = eager.linearize(T)
    assert el == ee and np.array_equal(Hl, He) and np.array_equal(bl, be)
    cl, sl = lazy.target_cov_state()
    ce, se = eager.target_cov_state()
    corr, _ = lazy.correspondences()
    used = np.unique(corr[corr >= 0])
    assert se.all() and 0 < sl.sum() == len(used) < len(tgt)            # only the correspondences were computed
    assert np.array_equal(np.nonzero(sl)[0], used)
    assert np.array_equal(cl[used], ce[used])                          # bit for bit
    # a second pose: new correspondences are added, the old ones kept
    T2 = T.copy()
    T2[:3, 3] += [0.3, 0.2, 0.0]
    assert lazy.linearize(T2)[0] == eager.linearize(T2)[0]
    cl2, sl2 = lazy.target_cov_state()
    assert sl2.sum() > sl.sum() and np.array_equal(cl2[sl2 > 0], ce[sl2 > 0])
    # whole aligns (look-ahead linearizes included) are bit-identical, final Hessian too
    Tl, Te = lazy.align(), eager.align()
    assert np.array_equal(Tl, Te) and np.array_equal(lazy.getFinalHessian(), eager.getFinalHessian())
    assert lazy.last_result == {**eager.last_result, "device_ms": lazy.last_result["device_ms"]}
    # consumers that need ALL covariances still get them: identical to the eager ones
    assert np.array_equal(lazy.getTargetCovariances(), eager.getTargetCovariances())
    # ... and after that the object simply uses the complete set
    assert lazy.linearize(T)[0] == ee
    # swap: the (partially covered) target becomes the source and needs every covariance
    lazy2, eager2 = make(True), make(False)
    lazy2.linearize(T)
    eager2.linearize(T)
    lazy2.swapSourceAndTarget()
    eager2.swapSourceAndTarget()
    assert np.array_equal(lazy2.align(), eager2.align())


def test_on_demand_keeps_the_parameters_of_the_first_align(rgc, orc, small_pair):
    """fast_gicp_impl.hpp:104-109: covariances are computed once per cloud, with the k / regularisation in
    force at that moment; changing k later does not touch them.  On-demand mode latches the same way."""
    src, tgt, _ = small_pair
    res = []
    for on_demand in (True, False):
        g = rgc.FastGICP()
        g.setTargetCovarianceMode(on_demand)
        g.setCorrespondenceRandomness(10)
        g.setInputTarget(tgt)
        g.setInputSource(src)
        g.align()
        g.setCorrespondenceRandomness(25)   # after the first align: ignored by the cached covariances
        T = np.eye(4)
        T[:3, 3] = [0.5, 0.3, 0.0]          # new correspondences -> new on-demand covariances, still k = 10
        res.append(g.linearize(T))
    assert res[0][0] == res[1][0] and np.array_equal(res[0][1], res[1][1])


def test_non_finite_input_is_rejected(rgc, small_pair):
    src, tgt, _ = small_pair
    for bad in (np.nan, np.inf, -np.inf):
        for which in ("source", "target"):
            g = rgc.FastGICP()
            cloud = (src if which == "source" else tgt).copy()
            cloud[len(cloud) // 3, 1] = bad
            with pytest.raises(rgc.RgcError, match="non-finite"):
                getattr(g, "setInputSource" if which == "source" else "setInputTarget")(cloud)
                g.waitInputs()  # the bounding box of a deferred build is examined by the next call
    raw = np.concatenate([src[:, :3], np.zeros((len(src), 1), np.float32)], 1)
    raw[5, 0] = np.nan
    with pytest.raises(rgc.RgcError, match="non-finite"):
        rgc.voxel_grid(raw, 0.3)
    # the pool is intact afterwards (scratch blocks returned on the failure path)
    g = rgc.FastGICP()
    g.setInputTarget(tgt)
    g.setInputSource(src)
    g.align()
    assert g.hasConverged()


@pytest.mark.parametrize("search", ["DIRECT7", "DIRECT27", "DIRECT1"])
def test_vgicp_fitness_between_aligns(rgc, orc, scan_pair, search):
    """align -> getFitnessScore -> align on FastVGICP used to shrink the shared
    partial-sum buffer under the voxel kernels (out-of-bounds write for DIRECT7 / DIRECT27)."""
    src, tgt, _ = scan_pair
    g = rgc.FastVGICP()
    g.setNeighborSearchMethod(getattr(rgc, search))
    g.setInputTarget(tgt)
    g.setInputSource(src)
    T1 = g.align()
    f1 = g.getFitnessScore()
    T2 = g.align()
    f2 = g.getFitnessScore()
    e1 = g.evaluateCost(T2)
    assert np.array_equal(T1, T2) and f1 == f2
    o = orc.FastVGICP(search_method=getattr(orc, search))
    o.setInputTarget(tgt)
    o.setInputSource(src)
    To = o.align()
    assert np.abs(T1[:3, 3] - To[:3, 3]).max() < 1e-4 and rot_angle(T1[:3, :3], To[:3, :3]) < 1e-5
    assert abs(f1 - o.getFitnessScore()) <= 1e-6 * f1
    assert np.isfinite(e1)


def test_vgicp_additive_weighted_is_additive(rgc, orc, small_pair):
    """VoxelAccumulationMode::ADDITIVE_WEIGHTED builds AdditiveGaussianVoxel like ADDITIVE
    (fast_vgicp_voxel.hpp:138-141)."""
    src, tgt, _ = small_pair
    out = []
    for mode in (rgc.ADDITIVE_WEIGHTED, rgc.ADDITIVE):
        g = rgc.FastVGICP()
        g.setVoxelAccumulationMode(mode)
        g.setInputTarget(tgt)
        g.setInputSource(src)
        out.append((g.align(), g.voxels()))
    assert np.array_equal(out[0][0], out[1][0])
    for a, b in zip(out[0][1], out[1][1]):
        assert np.array_equal(a, b)
    o = orc.FastVGICP(voxel_mode=orc.ADDITIVE_WEIGHTED)
    o.setInputTarget(tgt)
    o.setInputSource(src)
    To = o.align()
    assert np.abs(out[0][0][:3, 3] - To[:3, 3]).max() < 1e-4 and rot_angle(out[0][0][:3, :3], To[:3, :3]) < 1e-5


def test_c4_pairs_match_oracle(rgc, orc):
    """Config C4 inputs: VLP-16 sweep vs 100 000-point submap, initial error U(+-0.5 m, +-5 deg)
    about the truth, loop-closure parameters — 32 pairs against the oracle at 1e-4 m / 1e-5 rad, plus the
    acceptance gate of the caller (hasConverged && fitness <= 0.1... RGC_mapping.cpp:2070-2071)."""
    from rgc_slam_b200 import workloads
    pairs = workloads.make_c4_pairs(0, 32, n_base=4)
    oracles = {}
    worst_t = worst_r = 0.0
    accepted = 0
    for p in pairs:
        if p["base"] not in oracles:  # the oracle keeps its target kd-tree + covariances per submap, like the reference
            o = orc.FastGICP(max_iterations=64, corr_dist=2.0)
            o.setInputTarget(p["tgt"])
            o.setInputSource(p["src"])
            oracles[p["base"]] = o
        o = oracles[p["base"]]
        To = o.align(p["guess"])
        g = rgc.FastGICP()
        g.setMaximumIterations(64)
        g.setMaxCorrespondenceDistance(2.0)
        g.setInputTarget(p["tgt"])
        g.setInputSource(p["src"])
        T = g.align(p["guess"])
        _same_run(g, o, T, To)
        worst_t = max(worst_t, float(np.abs(T[:3, 3] - To[:3, 3]).max()))
        worst_r = max(worst_r, rot_angle(T[:3, :3], To[:3, :3]))
        fg, fo = g.getFitnessScore(), o.getFitnessScore()
        assert abs(fg - fo) <= 1e-6 * fo
        assert (g.hasConverged() and fg <= 0.1) == (o.last["converged"] and fo <= 0.1)
        accepted += int(g.hasConverged() and fg <= 0.1)
        E = np.linalg.inv(p["truth"]) @ T.astype(np.float64)
        assert np.linalg.norm(E[:3, 3]) < 0.05  # and it is the right answer
    assert accepted >= 30
    print(f"C4 parity: 32 pairs, worst |dt| {worst_t:.2e} m, worst angle {worst_r:.2e} rad, accepted {accepted}")


@pytest.mark.parametrize("world", [2, 3])
def test_c5_downscaled_fake_shards_match_oracle(rgc, orc, world):
    """Config C5 down-scaled (2 M-point map = 4 tiles of the 500k submap, 64-beam sweep, corr 2 m) on ONE GPU:
    `world` registration objects each hold one slab (+ halo) of the map and own the source points whose
    transformed position falls into it; the summed partial (err, H, b) and the union of the
    correspondences must equal the ORACLE's unsharded linearize."""
    import ctypes as C
    from rgc_slam_b200 import sharded, workloads
    case = workloads.make_c5_case(4, n_beams=64)
    src, tgt = case["src"], case["tgt"]
    assert len(tgt) == 2_000_000
    corr, cov_halo = 2.0, 4.0
    T = case["guess"].astype(np.float64)
    o = orc.FastGICP(corr_dist=corr)
    o.setInputTarget(tgt)
    o.setInputSource(src)
    eo, Ho, bo = o.linearize(T)
    oc, od2 = o.correspondences()
    L = rgc.lib()
    L.rgc_reg_set_owner_slab.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
    edges = sharded.slab_boundaries(tgt[:, 0], world)
    es, Hs, bs = 0.0, np.zeros((6, 6)), np.zeros(6)
    merged = np.full(len(src), -1, np.int64)
    big = float(np.finfo(np.float32).max)
    for r in range(world):
        g = rgc.FastGICP()
        g.setMaxCorrespondenceDistance(corr)
        keep = sharded.slab_select(tgt, 0, edges[r], edges[r + 1], corr + cov_halo)
        g.setInputTarget(np.ascontiguousarray(tgt[keep]))
        g.setInputSource(src)
        g.ctx.check(L.rgc_reg_set_owner_slab(g._h, 0, max(float(edges[r]), -big), min(float(edges[r + 1]), big)))
        er, Hr, br = g.linearize(T)
        es, Hs, bs = es + er, Hs + Hr, bs + br
        c, _ = g.correspondences()
        own = c >= 0
        assert (merged[own] == -1).all()      # every source point is handled by exactly one shard
        merged[own] = keep[c[own]]
        g = None
    assert np.array_equal(merged, oc.astype(np.int64))
    assert abs(es - eo) <= 1e-9 * abs(eo)
    assert np.abs(Hs - Ho).max() <= 1e-9 * np.abs(Ho).max() and np.abs(bs - bo).max() <= 1e-9 * np.abs(bo).max()


def test_features_64_ring_branch(scene, traj):
    """scanRegistration.cpp:160-170 (HDL-64 ring formula, rings > 50 dropped): the N_SCANS == 64 branch."""
    from oracle import oracle as orc
    from rgc_slam_b200 import synth
    from rgc_slam_b200.features import extract_features
    from test_gpu_features import _compare
    scans = [synth.lidar_scan(scene, traj[f], n_beams=64, n_azimuth=450, seed=700 + f) for f in (3, 9, 17)]
    res, _ = extract_features(scans, n_rings=64)
    for b, s in enumerate(scans):
        o = orc.extract_features(s, n_scans=64)
        _compare(res[b], o, f"64-ring scan {b}")
        assert o["cloud_size"] > 15000 and len(o["corner_sharp"]) > 500
        rings = np.floor(o["cloud"][:, 3]).astype(int)
        assert rings.max() == 50 and rings.min() == 0


def test_python_force_rebuilds_a_refilled_buffer(rgc, small_pair):
    """identity caching keys on the Python object; `force=True` is the way to say the contents changed."""
    src, tgt, _ = small_pair
    buf = tgt.copy()
    g = rgc.FastGICP()
    g.setInputTarget(buf)
    g.setInputSource(src)
    T1 = g.align()
    buf[:, 0] += 0.25                     # refilled in place
    g.setInputTarget(buf)                 # same object: cached, like the same shared_ptr in the reference
    assert np.array_equal(g.align(), T1)
    g.setInputTarget(buf, force=True)
    T2 = g.align()
    assert abs((T2[0, 3] - T1[0, 3]) - 0.25) < 2e-3


@pytest.mark.gpu
def test_speculative_grid_geometry_hits_and_misses(rgc, orc):
    """The build generates its Morton keys with the previous cloud's grid geometry before its own bounding box is
    known (rgc_gicp.cu: build_phase1 / build_phase3).  A cloud that does not fit that grid must be re-sorted
    (a miss), one that fits must not, and the neighbour lists are the oracle's either way."""
    import ctypes as C
    L = rgc.lib()
    L.rgc_debug_build_stats.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    from rgc_slam_b200 import api
    ctx = api.default_context(0)

    def stats():
        a, b = C.c_ulonglong(), C.c_ulonglong()
        ctx.check(L.rgc_debug_build_stats(ctx._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    rng = np.random.default_rng(5)

    def cloud(n, sigma):
        P = np.ones((n, 4), np.float32)
        P[:, :3] = rng.normal(0, sigma, (n, 3))
        return P

    small, small2, big = cloud(4000, 2.0), cloud(3000, 2.0), cloud(6000, 150.0)
    from rgc_slam_b200 import api
    ctx = api.default_context(0)
    for P in (small, small2, big, small, big):  # fits / fits / needs more bits (miss) / fits the larger grid / fits
        idx, d2 = rgc.knn(P, P[:500], 20)
        oi, od = orc.knn(P, P[:500], 20, brute=True)
        assert np.array_equal(idx, oi) and np.array_equal(d2, od)
    s0 = stats()
    rgc.knn(small, small[:10], 5)
    rgc.knn(cloud(5000, 400.0), small[:10], 5)  # far larger extent than anything before: must be redone
    s1 = stats()
    assert s1[0] >= s0[0] + 1, "speculative builds are not happening"
    assert s1[1] >= s0[1] + 1, "a cloud outside the hinted grid was not detected"


@pytest.mark.gpu
def test_speculative_table_sizes_hits_and_misses(rgc, orc):
    """The level tables of a cloud are allocated, cleared and filled with the table sizes of the previous cloud of the
    lane, before this cloud's own cell counts reach the host (rgc_gicp.cu: build_passes / build_phase3).  Similar
    clouds must keep those tables, a cloud with many more cells must get them rebuilt, and the neighbour lists are
    the oracle's either way (they do not depend on table sizes)."""
    import ctypes as C
    L = rgc.lib()
    L.rgc_debug_table_stats.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    from rgc_slam_b200 import api
    ctx = api.default_context(0)

    def stats():
        a, b = C.c_ulonglong(), C.c_ulonglong()
        ctx.check(L.rgc_debug_table_stats(ctx._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    rng = np.random.default_rng(11)

    def cloud(n):
        P = np.ones((n, 4), np.float32)
        P[:, :3] = rng.uniform(-20, 20, (n, 3))
        return P

    def check(P):
        idx, d2 = rgc.knn(P, P[:400], 20)
        oi, od = orc.knn(P, P[:400], 20, brute=True)
        assert np.array_equal(idx, oi) and np.array_equal(d2, od)

    check(cloud(3000))
    check(cloud(3000))   # (the first one may still shrink the hinted grid left by earlier tests: the sizes are kept from the second on)
    s0 = stats()
    check(cloud(3100))   # same extent, same density: the previous sizes hold
    check(cloud(2900))
    s1 = stats()
    assert s1[0] >= s0[0] + 2, "tables are not being filled speculatively"
    assert s1[1] == s0[1], "similar clouds must keep the speculative tables"
    check(cloud(40000))  # far more points than the cloud the sizes were made for: they are not even tried (a table that
    s2 = stats()         # runs full makes every further insert probe all of it)
    assert s2 == s1, "a much larger cloud must not be filled into the previous cloud's tables"

    def clustered(n):    # same extent (eight corner points), but nearly all points in a handful of cells
        P = np.ones((n, 4), np.float32)
        P[:, :3] = rng.normal(0, 0.003, (n, 3))
        P[:8, :3] = np.array([[sx, sy, sz] for sx in (-20, 20) for sy in (-20, 20) for sz in (-20, 20)], np.float32)
        return P

    check(clustered(3000))  # oversized tables are fine; the hint follows the latest cloud
    check(clustered(3000))
    s3 = stats()
    check(cloud(3000))   # as many points, > 10x the cells: far over the 70 % load the check tolerates (levels overflow outright)
    s4 = stats()
    assert s4[0] >= s3[0] + 1 and s4[1] >= s3[1] + 1, "an overfull speculative table was not detected"
    check(cloud(3000))
    check(cloud(3000))


@pytest.mark.gpu
@pytest.mark.parametrize("k", [33, 64, 128])
def test_k_above_32(rgc, orc, small_pair, k):
    """setCorrespondenceRandomness accepts any k in the reference (fast_gicp_impl.hpp:38-40); above 32 the tile kernel
    alone does the self-kNN, the generic covariance kernel follows and the target covariances are computed eagerly."""
    src, tgt, _ = small_pair
    idx = rgc.knn_self(tgt, k)
    oi, _ = orc.knn(tgt, tgt, k)
    assert np.array_equal(idx, oi)
    qi, qd = rgc.knn(tgt, src[:700], k)
    oq, od = orc.knn(tgt, src[:700], k)
    assert np.array_equal(qi, oq) and np.array_equal(qd, od)
    g = rgc.FastGICP()
    g.setCorrespondenceRandomness(k)
    g.setMaxCorrespondenceDistance(2.0)
    g.setInputSource(src)
    g.setInputTarget(tgt)
    c = g.getTargetCovariances()
    oc = orc.covariances_from_knn(tgt, oi, 3)
    scale = np.abs(oc).max(axis=(1, 2), keepdims=True)
    assert (np.abs(c - oc) / scale).max() < 1e-8
    if k == 64:
        T = g.align()
        o = orc.FastGICP(k=k, corr_dist=2.0)
        o.setInputSource(src)
        o.setInputTarget(tgt)
        To = o.align()
        _same_run(g, o, T, To)


def test_programmatic_dependent_launch_changes_nothing(rgc, scan_pair):
    """The kernels of the build and LM chains start with `griddepcontrol.wait` (pdl_enter, rgc_common.cuh) and are
    launched with programmatic stream serialization (launch_pdl, rgc_gicp.cu): a grid's blocks may become resident
    while its predecessor drains, but run only once it has completed.  RGC_NO_PDL=1 (read when a context is created)
    launches the same kernels plainly; pose, Hessian, iteration counts and neighbour lists must not differ by a bit."""
    src, tgt, _ = scan_pair

    def run(ctx):
        g = rgc.FastGICP(ctx)
        g.setInputTarget(tgt)
        g.setInputSource(src)
        T = g.align(np.eye(4, dtype=np.float32))
        r = g.last_result
        out = (T.copy(), g.getFinalHessian().copy(), (r["iterations"], r["n_linearize"], r["n_compute_error"]), g.getFitnessScore())
        g = None
        return out

    ctx_pdl = rgc.Context(0)
    os.environ["RGC_NO_PDL"] = "1"
    try:
        ctx_plain = rgc.Context(0)
    finally:
        del os.environ["RGC_NO_PDL"]
    a, b = run(ctx_pdl), run(ctx_plain)
    ctx_pdl.close()
    ctx_plain.close()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2] and a[3] == b[3]
