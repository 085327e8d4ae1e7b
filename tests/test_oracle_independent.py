"""CPU: the oracle against a SECOND, independent restatement of the reference, written in numpy in the
reference's own homogeneous 4x4 form (the C++ oracle works on 3x3 blocks with hand-rolled algebra; this file
uses numpy's 4x4 inverse, matrix products and scipy's kd-tree, i.e. none of the oracle's code).  The reference
cannot be built here and holds no golden vectors, so agreement of two restatements written
from the same source lines is the strongest pin available:

  update_correspondences   fast_gicp_impl.hpp:115-151   (float transform, 1-NN, threshold, RCR(3,3)=1, inverse, (3,3)=0)
  linearize                fast_gicp_impl.hpp:155-211   (4x6 Jacobian [skew(T a) | -I], H = J^T M J, b = J^T M e)
  compute_error            fast_gicp_impl.hpp:214-237   (stale correspondences and Mahalanobis matrices)
  step_lm                  lsq_registration_impl.hpp:125-172
  per-point feature pass   scanRegistration.cpp:233-294
"""
import numpy as np
import pytest
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation

from oracle import oracle as orc
from rgc_slam_b200 import synth


def so3_exp(w):
    """so3.hpp:58-77 is the rotation by |w| about w (tests/test_oracle_linalg.py checks the oracle's against this too)"""
    return Rotation.from_rotvec(np.asarray(w, np.float64)).as_matrix()


def skewd(v):
    """so3.hpp:21-27"""
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], np.float64)


class NumpyGICP:
    """fast_gicp_impl.hpp, statement by statement, on 4x4 / 4-vector numpy arrays"""

    def __init__(self, src, tgt, cov_src, cov_tgt, corr_dist=np.finfo(np.float32).max):
        self.src, self.tgt = src.astype(np.float32), tgt.astype(np.float32)   # xyz1 rows (getVector4fMap)
        self.cov_src, self.cov_tgt = cov_src, cov_tgt                          # Matrix4d per point
        self.tree = cKDTree(self.tgt[:, :3].astype(np.float64))
        self.thr = np.float32(corr_dist)

    def update_correspondences(self, T):
        Tf = T.astype(np.float32)                                              # :119
        q = (self.src @ Tf.T).astype(np.float32)                               # :131 (float product)
        d, idx = self.tree.query(q[:, :3].astype(np.float64), k=1)
        # the tree only proposes the neighbour; the distance is recomputed in float like flann::L2_Simple
        diff = q[:, :3] - self.tgt[idx, :3]
        sq = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]
        with np.errstate(over="ignore"):
            thr2 = self.thr * self.thr                                         # float product: +inf for FLT_MAX (:136)
        self.corr = np.where(sq < thr2, idx, -1)
        self.maha = np.zeros((len(self.src), 4, 4))
        for i in np.nonzero(self.corr >= 0)[0]:
            RCR = self.cov_tgt[self.corr[i]] + T @ self.cov_src[i] @ T.T        # :146
            RCR[3, 3] = 1.0                                                    # :147
            M = np.linalg.inv(RCR)                                             # :149
            M[3, 3] = 0.0                                                      # :150
            self.maha[i] = M

    def linearize(self, T):
        self.update_correspondences(T)
        err, H, b = 0.0, np.zeros((6, 6)), np.zeros(6)
        for i in np.nonzero(self.corr >= 0)[0]:
            mean_A = self.src[i].astype(np.float64)                            # :173
            mean_B = self.tgt[self.corr[i]].astype(np.float64)                 # :176
            tA = T @ mean_A                                                    # :179
            e = mean_B - tA                                                    # :180
            err += e @ self.maha[i] @ e                                        # :182
            J = np.zeros((4, 6))                                               # :188-190
            J[:3, :3] = skewd(tA[:3])
            J[:3, 3:] = -np.eye(3)
            H += J.T @ self.maha[i] @ J                                        # :194
            b += J.T @ self.maha[i] @ e                                        # :195
        return err, H, b

    def compute_error(self, T):
        err = 0.0
        for i in np.nonzero(self.corr >= 0)[0]:
            e = self.tgt[self.corr[i]].astype(np.float64) - T @ self.src[i].astype(np.float64)
            err += e @ self.maha[i] @ e
        return err


@pytest.fixture(scope="module")
def tiny_pair(scene, traj):
    tgt = synth.to_xyz1(synth.lidar_scan(scene, traj[10], n_azimuth=225, seed=5))
    src = synth.to_xyz1(synth.lidar_scan(scene, traj[11], n_azimuth=225, seed=6))
    return src, tgt


@pytest.mark.parametrize("corr_dist", [np.finfo(np.float32).max, 1.0])
def test_linearize_and_compute_error_equal_the_4x4_numpy_restatement(tiny_pair, corr_dist):
    src, tgt = tiny_pair
    o = orc.FastGICP(corr_dist=corr_dist)
    o.setInputTarget(tgt)
    o.setInputSource(src)
    T = np.eye(4)
    T[:3, :3] = so3_exp(np.array([0.01, -0.02, 0.03]))
    T[:3, 3] = [0.12, -0.05, 0.02]
    e, H, b = o.linearize(T)                      # also computes the covariances (lazily, like align)
    n = NumpyGICP(src, tgt, o.getSourceCovariances(), o.getTargetCovariances(), corr_dist)
    en, Hn, bn = n.linearize(T)
    assert np.array_equal(o.correspondences()[0], n.corr)
    if corr_dist == 1.0:
        assert (n.corr < 0).any() and (n.corr >= 0).any()   # the threshold is exercised both ways
    assert abs(e - en) <= 1e-10 * abs(en)
    assert np.abs(H - Hn).max() <= 1e-10 * np.abs(Hn).max()
    assert np.abs(b - bn).max() <= 1e-10 * np.abs(bn).max()
    # compute_error at ANOTHER pose keeps the correspondences and Mahalanobis matrices of the linearize (:214-237)
    T2 = T.copy()
    T2[:3, 3] += [0.03, 0.01, -0.02]
    assert abs(o.compute_error(T2) - n.compute_error(T2)) <= 1e-10 * abs(n.compute_error(T2))


def numpy_lm(n, guess, max_it):
    """LsqRegistration::computeTransformation + step_lm + is_converged (lsq_registration_impl.hpp:53-172) over a NumpyGICP"""
    x0 = guess.astype(np.float64)                                              # :54
    lam, lam_factor, max_inner = -1.0, 1e-9, 10                                # :17-19, :56
    rot_eps, trans_eps = 2e-3, 5e-4                                            # :12-13
    converged, iters, final_H, n_lin, n_ce = False, 0, np.eye(6), 0, 0

    def is_converged(D):                                                       # :81-91
        return max(np.abs(D[:3, :3] - np.eye(3)).max() / rot_eps, np.abs(D[:3, 3]).max() / trans_eps) < 1

    for it in range(max_it):                                                   # :65
        iters = it                                                             # nr_iterations_ = i (:66)
        y0, H, b = n.linearize(x0)                                             # :128
        n_lin += 1
        if lam < 0.0:
            lam = lam_factor * np.abs(np.diag(H)).max()                        # :130-132
        nu, ok, delta = 2.0, False, None
        for _ in range(max_inner):                                             # :135
            d = np.linalg.solve(H + lam * np.eye(6), -b)                       # :136-137 (LDLT there)
            D = np.eye(4)
            D[:3, :3] = so3_exp(d[:3])                                         # :139-141
            D[:3, 3] = d[3:]
            xi = D @ x0                                                        # :143 (left-multiplied)
            yi = n.compute_error(xi)                                           # :144
            n_ce += 1
            rho = (y0 - yi) / (d @ (lam * d - b))                              # :145
            if rho < 0:                                                        # :155-163
                if is_converged(D):
                    ok, delta = True, D
                    break
                lam, nu = nu * lam, 2 * nu
                continue
            x0 = xi                                                            # :165-168
            lam = lam * max(1.0 / 3.0, 1 - (2 * rho - 1) ** 3)
            final_H, ok, delta = H, True, D
            break
        if not ok:
            break                                                              # "lm not converged!!" (:69-72)
        if is_converged(delta):                                                # :74
            converged = True
            break
    return x0, iters, converged, final_H, n_lin, n_ce


def far_guess(trial: int) -> np.ndarray:
    """the same draws as tests/test_gpu_round2.py::far_guess (rotation U(20, 180) deg, translation U(+-5 m)^3)"""
    rng = np.random.default_rng(0)
    for _ in range(trial + 1):
        ang = rng.uniform(20, 180)
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        g = np.eye(4, dtype=np.float32)
        g[:3, :3] = Rotation.from_rotvec(np.deg2rad(ang) * axis).as_matrix()
        g[:3, 3] = rng.uniform(-5, 5, 3)
    return g


@pytest.mark.parametrize("case", ["near", "far"])
def test_lm_steps_equal_the_numpy_restatement(tiny_pair, case):
    """lsq_registration_impl.hpp:53-172 driven by the numpy linearize / compute_error above: the same number of
    linearize / compute_error calls (i.e. the same accepted and REJECTED steps: the far guess is one of the cases the
    GPU suite uses for the rho < 0 branch), iteration count, final pose and final Hessian as the oracle's align()."""
    src, tgt = tiny_pair
    if case == "near":
        max_it, corr_dist = 8, np.finfo(np.float32).max
        guess = np.eye(4, dtype=np.float32)
        guess[:3, 3] = [0.2, 0.1, 0.0]
    else:
        max_it, corr_dist, guess = 30, 1.0, far_guess(3)
    o = orc.FastGICP(max_iterations=max_it, corr_dist=corr_dist)
    o.setInputTarget(tgt)
    o.setInputSource(src)
    o.linearize(np.eye(4), want_Hb=False)
    n = NumpyGICP(src, tgt, o.getSourceCovariances(), o.getTargetCovariances(), corr_dist)
    To = o.align(guess)

    x0, iters, converged, final_H, n_lin, n_ce = numpy_lm(n, guess, max_it)
    assert (n_lin, n_ce) == (o.last["n_linearize"], o.last["n_compute_error"])
    if case == "far":
        assert n_ce > n_lin, "this case is here for its rejected steps"
    assert converged == o.last["converged"] and iters == o.last["iterations"]
    assert np.abs(x0 - To.astype(np.float64)).max() < 2e-6                      # float final_transformation_
    assert np.abs(final_H - o.last["final_hessian"]).max() <= 1e-7 * np.abs(final_H).max()


def test_feature_per_point_pass_equals_numpy(scene, traj):
    """scanRegistration.cpp:272-294 on the ring-ordered cloud the oracle emits: curvature from the +-5 stencil in
    float (left-to-right sums, `- 10 * x`), scaled by the double expression 2 / (1 + r / 20); range curvature in
    double from `- 10.0 * range`."""
    scan = synth.lidar_scan(scene, traj[12], seed=31)
    f = orc.extract_features(scan, n_scans=16)
    P = f["cloud"][:, :3].astype(np.float32)
    m = f["cloud_size"]
    rng = np.sqrt((P[:, 0] * P[:, 0] + P[:, 1] * P[:, 1]) + P[:, 2] * P[:, 2]).astype(np.float32)   # :237 (float sqrt of a float sum)
    idx = np.arange(5, m - 5)
    d = np.zeros((len(idx), 3), np.float32)
    for o_ in (-5, -4, -3, -2, -1):                                            # :272-274: p[i-5] + ... + p[i-1] - 10 p[i] + p[i+1] ...
        d = (d + P[idx + o_]).astype(np.float32)
    d = (d - np.float32(10) * P[idx]).astype(np.float32)
    for o_ in (1, 2, 3, 4, 5):
        d = (d + P[idx + o_]).astype(np.float32)
    dis = (2.0 / (1.0 + rng[idx].astype(np.float64) / 20.0))                    # :277 (double)
    sq = ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(np.float32)
    curv = (sq.astype(np.float64) * dis).astype(np.float32)                    # :279
    got = f["curvature"][idx]
    # the summation order of the stencil inside the reference expression is restated by the oracle term by term;
    # here it is grouped differently, so agreement is to float round-off of the stencil, not bit for bit
    assert np.allclose(got, curv, rtol=2e-4, atol=1e-6)
    r64 = rng.astype(np.float64)
    dr = np.zeros(len(idx))
    for o_ in (-5, -4, -3, -2, -1, 1, 2, 3, 4, 5):
        dr += r64[idx + o_]
    dr -= 10.0 * r64[idx]                                                      # :293
    c2 = np.abs(dr * dis).astype(np.float32)                                   # :294
    assert np.allclose(f["curvature2"][idx], c2, rtol=2e-4, atol=2e-5)


# ------------------------------------------------------------------------------------------------ FastVGICP
def numpy_voxelmap(tgt, covs, resolution, multiplicative):
    """fast_vgicp_voxel.hpp:129-160 (create_voxelmap) with the two voxel types of :83-127, in input order"""
    vox = {}
    for p, C in zip(tgt.astype(np.float32), covs):
        x = p.astype(np.float64)
        key = tuple(np.floor(x[:3] / resolution - 0.5).astype(int))               # voxel_coord :162-164
        v = vox.setdefault(key, dict(n=0, mean=np.zeros(4), cov=np.zeros((4, 4))))
        v["n"] += 1
        if multiplicative:                                                      # :93-101
            Ci = C.copy()
            Ci[3, 3] = 1
            Ci = np.linalg.inv(Ci)
            v["cov"] += Ci
            v["mean"] += Ci @ x
        else:                                                                   # :116-120
            v["mean"] += x
            v["cov"] += C
    for v in vox.values():
        if multiplicative:                                                      # :103-109
            v["cov"][3, 3] = 1
            v["mean"][3] = 1
            v["cov"] = np.linalg.inv(v["cov"])
            v["mean"] = v["cov"] @ v["mean"]
        else:                                                                   # :122-125
            v["mean"] /= v["n"]
            v["cov"] /= v["n"]
    return vox


def neighbor_offsets(method):
    """fast_vgicp_voxel.hpp:10-44"""
    if method == orc.DIRECT1:
        return [(0, 0, 0)]
    if method == orc.DIRECT7:
        return [(0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
    return [(i - 1, j - 1, k - 1) for i in range(3) for j in range(3) for k in range(3)]


def numpy_vgicp_linearize(src, cov_src, vox, resolution, method, T, T_err=None):
    """fast_vgicp_impl.hpp:73-160: voxel correspondences + Mahalanobis matrices at T, sums at T_err (= T for
    linearize; another pose for compute_error, which keeps the correspondences, :179-204)"""
    T_err = T if T_err is None else T_err
    err, H, b, ncorr = 0.0, np.zeros((6, 6)), np.zeros(6), 0
    for p, CA in zip(src.astype(np.float32), cov_src):
        a = p.astype(np.float64)
        coord = np.floor((T @ a)[:3] / resolution - 0.5).astype(int)            # :85-87
        for off in neighbor_offsets(method):
            v = vox.get(tuple(coord + np.array(off)))
            if v is None:
                continue
            ncorr += 1
            RCR = v["cov"] + T @ CA @ T.T                                       # :112-116
            RCR[3, 3] = 1.0
            M = np.linalg.inv(RCR)
            M[3, 3] = 0.0
            tA = T_err @ a
            e = v["mean"] - tA                                                  # :146-147
            w = np.sqrt(v["n"])                                                 # :149
            err += w * (e @ M @ e)
            J = np.zeros((4, 6))
            J[:3, :3] = skewd(tA[:3])
            J[:3, 3:] = -np.eye(3)
            H += w * (J.T @ M @ J)
            b += w * (J.T @ M @ e)
    return err, H, b, ncorr


@pytest.mark.parametrize("method", ["DIRECT1", "DIRECT7", "DIRECT27"])
@pytest.mark.parametrize("mode", ["ADDITIVE", "MULTIPLICATIVE"])
def test_vgicp_voxelmap_and_linearize_equal_the_numpy_restatement(tiny_pair, method, mode):
    src, tgt = tiny_pair
    src, tgt = src[::3], tgt                                                    # keep the pure-python loops short
    method_id, mode_id, res = getattr(orc, method), getattr(orc, mode), 1.0
    o = orc.FastVGICP(resolution=res, search_method=method_id, voxel_mode=mode_id)
    o.setInputTarget(tgt)
    o.setInputSource(src)
    T = np.eye(4)
    T[:3, :3] = so3_exp(np.array([0.01, 0.005, -0.02]))
    T[:3, 3] = [0.1, 0.05, -0.02]
    e, H, b = o.linearize(T)
    vox = numpy_voxelmap(tgt, o.getTargetCovariances(), res, mode == "MULTIPLICATIVE")
    coords, num, mean, cov = o.voxels()
    assert len(coords) == len(vox)
    # MULTIPLICATIVE sums inverses of covariances with condition number 1e3 and inverts the sum again: the two
    # restatements (cofactor inverse there, LAPACK here) agree to round-off times that conditioning
    tm, tc, th = (1e-11, 1e-9, 1e-9) if mode == "ADDITIVE" else (1e-8, 1e-6, 1e-7)
    for c, k, m, cv in zip(coords, num, mean, cov):
        v = vox[tuple(int(x) for x in c)]
        assert v["n"] == k
        assert np.allclose(v["mean"][:3], m, rtol=tm, atol=tm)
        full = np.array([[cv[0], cv[1], cv[2]], [cv[1], cv[3], cv[4]], [cv[2], cv[4], cv[5]]])
        assert np.allclose(v["cov"][:3, :3], full, rtol=tc, atol=tc * np.abs(full).max())
    en, Hn, bn, ncorr = numpy_vgicp_linearize(src, o.getSourceCovariances(), vox, res, method_id, T)
    assert ncorr == o.num_correspondences() and ncorr > 0
    assert abs(e - en) <= th * abs(en)
    assert np.abs(H - Hn).max() <= th * np.abs(Hn).max()
    assert np.abs(b - bn).max() <= th * np.abs(bn).max()
    T2 = T.copy()
    T2[:3, 3] += [0.02, -0.01, 0.01]
    en2 = numpy_vgicp_linearize(src, o.getSourceCovariances(), vox, res, method_id, T, T_err=T2)[0]
    assert abs(o.compute_error(T2) - en2) <= th * abs(en2)


# ------------------------------------------------------------------------------------------------ A-LOAM features
F32 = np.float32
GROUND_SCAN_RANGE = np.array([2.66, 3.04, 3.56, 4.30, 5.44, 7.41, 11.63, 27.12], np.float32)   # scanRegistration.cpp:40


def features_restated(cloud, inten_raw, scan_start, scan_end, n_scans, use_intensity=1):
    """scanRegistration.cpp:233-663 on the ring-ordered cloud (the output of the ring bucketing, :135-230), written
    from the reference text with numpy float32 / float64 / Python-int arithmetic where the C++ expression is float /
    double / int.  Ties of the two sorts: (key, index), the convention of SURVEY §8c."""
    P = cloud[:, :3].astype(F32)
    m = len(P)
    x, y, z = P[:, 0], P[:, 1], P[:, 2]
    rng = np.sqrt((x * x + y * y) + z * z)                                        # :235-238, float
    assert rng.dtype == F32
    sizes = [int(scan_end[i] + 5 - (scan_start[i] - 5)) for i in range(n_scans)]  # laserCloudScans[i].size() (:220-228)
    idx = np.arange(5, m - 5)

    scan_angle = np.zeros(m, F32)
    for i in idx[rng[idx] < 2]:                                                   # :240-256
        a, b, now = P[i + 5].astype(np.float64), P[i - 5].astype(np.float64), P[i].astype(np.float64)
        c = (a + b) / 2
        nrm = np.cross(a - b, now - c)
        sa = F32(nrm @ now / (np.linalg.norm(nrm) * np.linalg.norm(now)))
        scan_angle[i] = -sa if sa < 0 else sa
    near = (scan_angle.astype(np.float64) < 0.07) & (rng < 2)

    inten2 = [int(v) for v in inten_raw]                                          # deque<int> (:151, :231)
    inten = list(inten2)
    for i in idx[near[idx]]:                                                      # :258-269: every assignment truncates to int
        v = int(0.9 * inten2[i])
        for j in range(-5, 6):
            if j != 0:
                v = int(v + 0.005 * inten2[i + j])
        inten[i] = v
    inten_a = np.array(inten, np.int64)

    def stencil(v):                                                               # :272-274, float, left to right
        s = v[idx - 5]
        for o_ in (-4, -3, -2, -1):
            s = s + v[idx + o_]
        s = s - F32(10) * v[idx]
        for o_ in (1, 2, 3, 4, 5):
            s = s + v[idx + o_]
        assert s.dtype == F32
        return s

    dX, dY, dZ = stencil(x), stencil(y), stencil(z)
    dI = np.zeros(len(idx), np.int64)                                             # :275, int arithmetic
    for o_ in range(-5, 6):
        dI += inten_a[idx + o_] * (-10 if o_ == 0 else 1)
    dI = dI.astype(F32)
    dis = (2.0 / (1.0 + rng[idx].astype(np.float64) / 20.0)).astype(F32)          # :277
    dis[dis.astype(np.float64) < 0.2] = F32(0.2)                                  # :278
    curvature, inten_curv, curv2 = np.zeros(m, F32), np.zeros(m, F32), np.zeros(m, F32)
    dist_src, other_src = np.zeros(m, F32), np.zeros(m, F32)
    curvature[idx] = ((dX * dX + dY * dY) + dZ * dZ) * dis                        # :279
    dist_src[idx] = (0.5 + dis.astype(np.float64)).astype(F32)                    # :280
    sa = scan_angle[idx]
    other_src[idx] = np.where(near[idx], ((sa * F32(10)).astype(np.float64) + 0.6).astype(F32), F32(3))          # :282-293
    inten_curv[idx] = np.where(near[idx], ((sa.astype(np.float64) + 0.3) * dI.astype(np.float64)).astype(F32), dI)
    r = rng
    s5 = (((r[idx - 5] + r[idx - 4]) + r[idx - 3]) + r[idx - 2]) + r[idx - 1]     # float until `- 10.0 *` (:293)
    d = s5.astype(np.float64) - 10.0 * r[idx].astype(np.float64)
    for o_ in (1, 2, 3, 4, 5):
        d = d + r[idx + o_].astype(np.float64)
    curv2[idx] = np.abs(d.astype(F32) * dis)                                      # :294

    # ---- ground marking + plane (:307-431)
    marked = np.zeros(m, np.int32)
    ground_points, nearg, lw = [], [], []
    center, weights, start = np.zeros(3), 0.0, 0
    for i in range(7):                                                            # groundScanInd (:36)
        th = F32(0.8 * (1.0 + i // 6))                                            # int division (:323)
        gw = 1.5 - i // 6                                                         # :325
        for col in range(5, sizes[i] - 5):
            ci = start + col
            if abs(rng[ci] - GROUND_SCAN_RANGE[i]) < th and float(z[ci]) < 0.3:   # :324-330
                marked[ci] = 1
                for n in range(-5, 5):                                            # :333 (n < 5)
                    if abs(rng[ci + n] - rng[ci]) < th / F32(2):
                        marked[ci + n] = 1
                        ground_points.append(ci + n)
                        tmp = P[ci + n].astype(np.float64)
                        center = center + gw * tmp
                        weights += gw
                        nearg.append(tmp)
                        lw.append(gw)
        start += sizes[i]
    gp = np.zeros(11)
    if nearg:
        center = center / weights
        cov = np.zeros((3, 3))
        for p, w in zip(nearg, lw):
            t = p - center
            cov = cov + w * np.outer(t, t)
        cov = cov / weights
        ev, V = np.linalg.eigh(cov)                                               # SelfAdjointEigenSolver (:371)
        n0 = V[:, 0] / np.linalg.norm(V[:, 0])
        if center @ n0 < 0:
            n0 = -n0
        g1, dist = 0.0, 0.0
        for p in nearg:                                                           # :386-400
            t = p - center
            nt = np.linalg.norm(t)
            t = t / nt if nt > 0 else t
            dw = 1 - 100 * abs(n0 @ t)
            if dw < 0:
                dw = 0.1
            g1 += dw
            dist += dw * (n0 @ p)
        dist = dist / g1
        g1 = g1 / len(nearg)
        lader_h = 0.56
        if dist / lader_h > 1.1 or dist / lader_h < 0.9:                           # :404-413
            dist = lader_h
        if g1 < 0.9:
            dist = 0.9 * lader_h + 0.1 * dist
        gp = np.concatenate([n0, V[:, 1], V[:, 2], [dist, 1 - g1]])

    # ---- occlusion (:433-456)
    picked = np.zeros(m + 8, np.int32)
    for i in idx:
        d1, d2 = rng[i], rng[i + 1]
        if float(d1 - d2) > 0.04 * float(d2):
            picked[i - 5:i + 1] = 1
        elif float(d2 - d1) > 0.04 * float(d1):
            picked[i + 1:i + 7] = 1

    # ---- sextant sort + greedy selection (:469-644)
    label, inten_label = np.zeros(m, np.int32), np.zeros(m, np.int32)
    inten_picked = np.zeros(m + 8, np.int32)
    corner_sharp, corner_less, surf_flat, less_flat, inten_sharp, inten_less = [], [], [], [], [], []

    def gap2(a, b):                                                               # float sum vs the double 0.05
        dx, dy, dz = P[a] - P[b]
        return float(dx * dx + dy * dy + dz * dz)

    def suppress(ind):                                                            # :517-534 / :566-583
        for l in range(1, 6):
            if gap2(ind + l, ind + l - 1) > 0.05:
                break
            picked[ind + l] = 1
        for l in range(-1, -6, -1):
            if gap2(ind + l, ind + l + 1) > 0.05:
                break
            picked[ind + l] = 1

    for i in range(n_scans):
        s, e = int(scan_start[i]), int(scan_end[i])
        if e - s < 10:
            continue
        for j in range(6):
            sp, ep = s + (e - s) * j // 6, s + (e - s) * (j + 1) // 6 - 1
            by_curv = sorted(range(sp, ep + 1), key=lambda k: (curvature[k], k))
            by_inten = sorted(range(sp, ep + 1), key=lambda k: (inten_curv[k], k))
            n_big = 0
            for ind in reversed(by_curv):                                         # :487-536
                if picked[ind] == 0 and marked[ind] != 1 and float(curvature[ind]) > 0.1 and float(curv2[ind]) > 0.3:
                    n_big += 1
                    if n_big <= 20:
                        label[ind] = 2
                        corner_sharp.append(ind)
                        corner_less.append(ind)
                    elif n_big <= 21:
                        label[ind] = 1
                        corner_less.append(ind)
                    else:
                        break
                    picked[ind] = 1
                    suppress(ind)
            n_small = 0
            for ind in by_curv:                                                   # :540-584
                if picked[ind] == 0 and float(curvature[ind]) < 0.3 and float(curv2[ind]) < 0.4:
                    n_small += 1
                    if n_small <= 40:
                        label[ind] = -1
                        surf_flat.append(ind)
                    else:
                        break
                    picked[ind] = 1
                    suppress(ind)
            less_flat += [k for k in range(sp, ep + 1) if label[k] <= 0]          # :586-592
            n_int = 0
            for ind in reversed(by_inten):                                        # :595-641
                if inten_picked[ind] == 0 and marked[ind] != 1 and float(inten_curv[ind]) > 65 and label[ind] not in (1, 2):
                    n_int += 1
                    if n_int <= 20:
                        inten_label[ind] = 2
                        inten_sharp.append(ind)
                        inten_less.append(ind)
                    elif n_int <= 21:
                        inten_label[ind] = 1
                        inten_less.append(ind)
                        corner_less.append(ind)
                    else:
                        break
                    inten_picked[ind] = 1
                    for l in range(1, 6):
                        if abs(inten[ind + l] - inten[ind + l - 1]) > 35:
                            break
                        inten_picked[ind + l] = 1
                    for l in range(-1, -6, -1):
                        if abs(inten[ind + l] - inten[ind + l + 1]) > 35:
                            break
                        inten_picked[ind + l] = 1
    merged = 0
    if use_intensity and surf_flat and len(corner_sharp) / len(surf_flat) < 0.3:  # :645-652
        merged = 1
    return dict(range_vec=rng, scan_angle=scan_angle, intensity_num=np.array(inten, np.int32), curvature=curvature, inten_curvature=inten_curv,
                curvature2=curv2, distance_source=dist_src, other_source=other_src, ground_marked=marked, ground_points=np.array(ground_points, np.int32),
                groundparam=gp, neighbor_picked=picked[:m], inten_neighbor_picked=inten_picked[:m], label=label, inten_label=inten_label,
                corner_sharp=np.array(corner_sharp, np.int32), corner_less_sharp=np.array(corner_less, np.int32), surf_flat=np.array(surf_flat, np.int32),
                surf_less_flat=np.array(less_flat, np.int32), inten_sharp=np.array(inten_sharp, np.int32), inten_less_sharp=np.array(inten_less, np.int32),
                inten_merged=merged)


@pytest.mark.parametrize("beams,az,frame", [(16, 1800, 12), (32, 900, 20), (16, 1800, 40)])
def test_feature_path_equals_the_python_restatement(scene, traj, beams, az, frame):
    """Labels, masks, feature lists and per-point values of the oracle (C++, oracle/orc_features.hpp) against the
    Python restatement above: integers and lists bit for bit, floats bit for bit where the expression is elementwise."""
    scan = synth.lidar_scan(scene, traj[frame], n_beams=beams, n_azimuth=az, seed=100 + frame)
    f = orc.extract_features(scan, n_scans=beams)
    inten_raw = scan[f["src_index"], 3]
    g = features_restated(f["cloud"], inten_raw, f["scan_start"], f["scan_end"], beams)
    for k in ("range_vec", "intensity_num", "curvature", "inten_curvature", "curvature2", "distance_source", "other_source", "ground_marked",
              "neighbor_picked", "inten_neighbor_picked", "label", "inten_label", "ground_points", "corner_sharp", "corner_less_sharp", "surf_flat",
              "surf_less_flat", "inten_sharp", "inten_less_sharp"):
        assert np.array_equal(f[k], g[k]), k
    assert np.allclose(f["scan_angle"], g["scan_angle"], rtol=1e-6, atol=1e-7)
    assert f["inten_merged"] == g["inten_merged"]
    assert (f["label"] == 2).sum() > 50 and (f["label"] == -1).sum() > 200                                    # the case is not vacuous
    if beams == 16:  # (the reference's Ground_scan_range table is the VLP-16's: a 32-beam scan marks no ground)
        assert f["ground_marked"].sum() > 100 and len(f["ground_points"]) > 100
    a, b = f["groundparam"], g["groundparam"]
    assert np.allclose(a[:3], b[:3], atol=1e-9) and np.allclose(a[9:], b[9:], atol=1e-9)
    for o_ in (3, 6):                                                             # eigenvectors 1, 2: defined up to sign
        assert min(np.abs(a[o_:o_ + 3] - b[o_:o_ + 3]).max(), np.abs(a[o_:o_ + 3] + b[o_:o_ + 3]).max()) < 1e-7


def rings_restated(scan, n_scans, min_range=0.5, max_range=80.0, scan_period=0.1):
    """scanRegistration.cpp:110-230 + :732-763: NaN removal, range gate, ring id from the elevation, relative time from the
    azimuth (startOri / endOri / halfPassed), ring buckets concatenated in ring order.  Returns (src_index, ring, relTime)."""
    P = scan.astype(F32)
    keep = []
    th1, th2 = F32(min_range), F32(max_range)
    for i, (x, y, z, _) in enumerate(P):                                           # removeClosedPointCloud (:732-763)
        if not (np.isfinite(x) and np.isfinite(y) and np.isfinite(z)):
            continue                                                              # pcl::removeNaNFromPointCloud (:112)
        dis = (x * x + y * y) + z * z
        if dis < th1 * th1 or dis > th2 * th2:
            continue
        if x < 0 and abs(float(y)) < 0.5:
            continue
        keep.append(i)
    Q = P[keep]
    n = len(Q)
    pi = np.pi
    start = F32(-np.arctan2(Q[0, 1], Q[0, 0]))                                    # :117-118 (float)
    end = F32(np.float64(F32(-np.arctan2(Q[n - 1, 1], Q[n - 1, 0]))) + 2 * pi)
    if float(end - start) > 3 * pi:                                               # :120-127
        end = F32(float(end) - 2 * pi)
    elif float(end - start) < pi:
        end = F32(float(end) + 2 * pi)
    half = False
    buckets = [[] for _ in range(n_scans)]
    for k in range(n):
        x, y, z = Q[k, 0], Q[k, 1], Q[k, 2]
        va = F32(float(np.arctan(z / np.sqrt(x * x + y * y))) * 180 / pi)          # :141
        if n_scans == 16:                                                         # :144-178
            sid = int(float(va + F32(15)) / 2 + 0.5)
        elif n_scans == 32:
            sid = int((float(va) + 92.0 / 3.0) * 3.0 / 4.0)
        else:
            sid = int((2 - float(va)) * 3.0 + 0.5) if float(va) >= -8.83 else n_scans // 2 + int((-8.83 - float(va)) * 2.0 + 0.5)
            if float(va) > 2 or float(va) < -24.33 or sid > 50:
                continue
        if sid > n_scans - 1 or sid < 0:
            continue
        ori = float(F32(-np.arctan2(y, x)))                                       # :186
        s, e = float(start), float(end)
        if not half:                                                              # :187-205 (float ori, double pi terms)
            if ori < s - pi / 2:
                ori = float(F32(ori + 2 * pi))
            elif ori > s + pi * 3 / 2:
                ori = float(F32(ori - 2 * pi))
            if float(F32(ori) - start) > pi:
                half = True
        else:
            ori = float(F32(ori + 2 * pi))
            if ori < e - pi * 3 / 2:
                ori = float(F32(ori + 2 * pi))
            elif ori > e + pi / 2:
                ori = float(F32(ori - 2 * pi))
        rel = F32(F32(ori) - start) / (end - start)                               # :208, float
        buckets[sid].append((keep[k], sid, float(rel)))
    flat = [t for b in buckets for t in b]
    return np.array([t[0] for t in flat]), np.array([t[1] for t in flat]), np.array([t[2] for t in flat]), scan_period


@pytest.mark.parametrize("beams,az", [(16, 1800), (32, 900)])
def test_ring_assignment_equals_the_python_restatement(scene, traj, beams, az):
    """which points survive, their ring and their order are integers and must be equal; the relative time goes through
    atan2 in float (libm there, numpy here) and is compared to 2e-6"""
    scan = synth.lidar_scan(scene, traj[25], n_beams=beams, n_azimuth=az, seed=77)
    scan[::997, 0] = np.nan                                                       # a few invalid returns
    f = orc.extract_features(scan, n_scans=beams)
    src_index, ring, rel, period = rings_restated(scan, beams)
    assert np.array_equal(f["src_index"], src_index)
    inten = f["cloud"][:, 3].astype(np.float64)
    assert np.abs(inten - (ring + period * rel)).max() < 2e-6
    assert rel.min() > -0.5 and rel.max() < 1.5
    sizes = np.bincount(ring, minlength=beams)
    ends = np.cumsum(sizes)
    assert np.array_equal(f["scan_start"], ends - sizes + 5) and np.array_equal(f["scan_end"], ends - 5)   # :221-228


@pytest.mark.parametrize("max_range", [np.finfo(np.float64).max, 0.05])
def test_fitness_score_equals_numpy(tiny_pair, max_range):
    """pcl::Registration::getFitnessScore(max_range): the source transformed by the final transformation (float),
    1-NN squared distance to the target, mean over the distances <= max_range (callers RGC_odometer.cpp:1010,
    RGC_mapping.cpp:2070)."""
    src, tgt = tiny_pair
    o = orc.FastGICP(max_iterations=5)
    o.setInputTarget(tgt)
    o.setInputSource(src)
    T = o.align()
    q = (src.astype(F32) @ T.astype(F32).T)[:, :3]
    d, _ = cKDTree(tgt[:, :3].astype(np.float64)).query(q.astype(np.float64), k=1)
    d2 = d * d
    sel = d2 <= max_range
    assert sel.any() and (max_range > 1 or not sel.all())
    assert abs(o.getFitnessScore(max_range) - d2[sel].mean()) <= 1e-5 * d2[sel].mean()   # float d2 there, double here


@pytest.mark.parametrize("beams", [16, 32])
def test_golden_feature_vectors_equal_the_python_chain(beams):
    """tests/golden/features_small.npz (what the GPU suite is checked against) from the raw scan through the two Python
    restatements above — ring assignment, then the feature pass — with no oracle code in between."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "features_small.npz"))
    scan = z[f"scan{beams}"]
    src_index, ring, rel, period = rings_restated(scan, beams)
    assert np.array_equal(src_index, z[f"src_index{beams}"])
    cloud = np.zeros((len(src_index), 4), F32)
    cloud[:, :3] = scan[src_index, :3]
    cloud[:, 3] = (ring + period * rel).astype(F32)
    sizes = np.bincount(ring, minlength=beams)
    ends = np.cumsum(sizes)
    g = features_restated(cloud, scan[src_index, 3], ends - sizes + 5, ends - 5, beams)
    for k in ("label", "inten_label", "neighbor_picked", "inten_neighbor_picked", "ground_marked", "curvature", "inten_curvature", "curvature2",
              "corner_sharp", "surf_flat", "inten_sharp", "corner_less_sharp"):
        assert np.array_equal(g[k], z[f"{k}{beams}"]), k
    a, b = z[f"groundparam{beams}"], g["groundparam"]
    assert np.allclose(a[:3], b[:3], atol=1e-9) and np.allclose(a[9:], b[9:], atol=1e-9)


def test_golden_gicp_vectors_equal_the_numpy_chain():
    """tests/golden/gicp_small.npz (what the GPU suite is checked against): covariances from the golden neighbour lists
    with numpy's SVD (fast_gicp_impl.hpp:256-293, PLANE), then linearize and the LM loop of the numpy restatement."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "gicp_small.npz"))
    src, tgt = z["src"], z["tgt"]

    idx_t, _ = orc.knn(tgt, tgt, 20)
    assert np.array_equal(idx_t, z["knn_idx"])
    idx_s, _ = orc.knn(src, src, 20)
    ks = cKDTree(src[:, :3].astype(np.float64)).query(src[:, :3].astype(np.float64), k=20)[1]
    assert (np.sort(ks, 1) == np.sort(idx_s, 1)).mean() > 0.999                   # the neighbour SETS, independently (ties aside)
    n = NumpyGICP(src, tgt, plane_covs_numpy(src, idx_s), plane_covs_numpy(tgt, idx_t))
    e, H, b = n.linearize(z["T_lin"].astype(np.float64))
    assert np.array_equal(n.corr, z["corr"])
    assert abs(e - z["lin_err"]) <= 1e-7 * abs(e)
    assert np.abs(H - z["lin_H"]).max() <= 1e-7 * np.abs(H).max() and np.abs(b - z["lin_b"]).max() <= 1e-7 * np.abs(b).max()
    x0, iters, converged, _, _, _ = numpy_lm(n, np.eye(4, dtype=np.float32), 64)
    assert iters == int(z["iterations"]) and converged
    assert np.abs(x0 - z["T_final"].astype(np.float64)).max() < 1e-6


def plane_covs_numpy(P, idx):
    """fast_gicp_impl.hpp:256-293 (PLANE) with numpy's SVD, from given neighbour lists"""
    out = np.zeros((len(P), 4, 4))
    for i, nb in enumerate(idx):
        X = P[nb].astype(np.float64)
        X = X - X.mean(axis=0)
        U, _, Vt = np.linalg.svd((X.T @ X / len(nb))[:3, :3])
        out[i, :3, :3] = U @ np.diag([1.0, 1.0, 1e-3]) @ Vt
    return out


class NumpyVGICP:
    """FastVGICP's linearize / compute_error (fast_vgicp_impl.hpp:73-204) with the interface numpy_lm drives"""

    def __init__(self, src, cov_src, vox, resolution, method):
        self.src, self.cov_src, self.vox, self.res, self.method = src, cov_src, vox, resolution, method
        self.T_lin = None

    def linearize(self, T):
        self.T_lin = T.copy()
        e, H, b, self.ncorr = numpy_vgicp_linearize(self.src, self.cov_src, self.vox, self.res, self.method, T)
        return e, H, b

    def compute_error(self, T):
        return numpy_vgicp_linearize(self.src, self.cov_src, self.vox, self.res, self.method, self.T_lin, T_err=T)[0]


def test_golden_vgicp_vectors_equal_the_numpy_chain():
    """tests/golden/vgicp_small.npz: voxel map, DIRECT1 / DIRECT7 linearize and the aligned pose from the numpy chain alone"""
    import os
    gold = os.path.join(os.path.dirname(__file__), "golden")
    z, v = np.load(os.path.join(gold, "gicp_small.npz")), np.load(os.path.join(gold, "vgicp_small.npz"))
    tgt = z["tgt"]
    cov_t = plane_covs_numpy(tgt, z["knn_idx"])
    vox = numpy_voxelmap(tgt, cov_t, 1.0, False)
    keys = sorted(vox)
    assert np.array_equal(np.array(keys), v["vox_coords"]) and np.array_equal([vox[k]["n"] for k in keys], v["vox_num"])
    assert np.allclose([vox[k]["mean"][:3] for k in keys], v["vox_mean"], rtol=0, atol=1e-9)
    got_cov = np.array([[vox[k]["cov"][0, 0], vox[k]["cov"][0, 1], vox[k]["cov"][0, 2], vox[k]["cov"][1, 1], vox[k]["cov"][1, 2], vox[k]["cov"][2, 2]] for k in keys])
    assert np.allclose(got_cov, v["vox_cov"], rtol=0, atol=1e-9)
    full_src = z["src"]
    cov_s = plane_covs_numpy(full_src, orc.knn(full_src, full_src, 20)[0])
    T = z["T_lin"].astype(np.float64)
    for name, method in (("d1", orc.DIRECT1), ("d7", orc.DIRECT7)):
        e, H, b, ncorr = numpy_vgicp_linearize(full_src, cov_s, vox, 1.0, method, T)
        assert ncorr == int(v[f"{name}_ncorr"])
        assert abs(e - v[f"{name}_err"]) <= 1e-7 * abs(e)
        assert np.abs(H - v[f"{name}_H"]).max() <= 1e-7 * np.abs(H).max() and np.abs(b - v[f"{name}_b"]).max() <= 1e-7 * np.abs(b).max()
    x0, iters, converged, _, _, _ = numpy_lm(NumpyVGICP(full_src, cov_s, vox, 1.0, orc.DIRECT1), np.eye(4, dtype=np.float32), 64)
    assert converged and iters == int(v["d1_iterations"])
    assert np.abs(x0 - v["d1_T"].astype(np.float64)).max() < 1e-6
