"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes driver for oracle/liboracle.so.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference
legs.  The product package (rgc_slam_b200) never imports this module.  PARITY UNPINNED: the
reference has no golden vectors and cannot be built here; see DESIGN.md.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

REG_NONE, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS = range(5)
OPT_GN, OPT_LM = 0, 1

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.orc_gicp_create.restype = C.c_void_p
        L.orc_gicp_destroy.argtypes = [C.c_void_p]
        L.orc_gicp_set_params.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
        L.orc_gicp_set_source.argtypes = [C.c_void_p, _f32p, C.c_int]
        L.orc_gicp_set_target.argtypes = [C.c_void_p, _f32p, C.c_int]
        L.orc_gicp_ensure_covariances.argtypes = [C.c_void_p]
        L.orc_gicp_set_source_covs.argtypes = [C.c_void_p, _f64p, C.c_int]
        L.orc_gicp_set_target_covs.argtypes = [C.c_void_p, _f64p, C.c_int]
        L.orc_gicp_get_source_covs.argtypes = [C.c_void_p, _f64p]
        L.orc_gicp_get_target_covs.argtypes = [C.c_void_p, _f64p]
        L.orc_gicp_linearize.argtypes = [C.c_void_p, _f64p, C.c_void_p, C.c_void_p]
        L.orc_gicp_linearize.restype = C.c_double
        L.orc_gicp_compute_error.argtypes = [C.c_void_p, _f64p]
        L.orc_gicp_compute_error.restype = C.c_double
        L.orc_gicp_get_correspondences.argtypes = [C.c_void_p, _i32p, _f32p]
        L.orc_gicp_align.argtypes = [C.c_void_p, _f32p, _f32p, C.c_void_p, _i32p, _f64p]
        L.orc_gicp_align.restype = C.c_double
        L.orc_gicp_fitness.argtypes = [C.c_void_p, C.c_double]
        L.orc_gicp_fitness.restype = C.c_double
        L.orc_knn.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_int, _i32p, _f32p, C.c_int]
        L.orc_knn_bruteforce.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_int, _i32p, _f32p]
        L.orc_covariances_from_knn.argtypes = [_f32p, C.c_int, _i32p, C.c_int, C.c_int, _f64p]
        L.orc_extract_features.restype = C.c_double
        L.orc_max_threads.restype = C.c_int
    return _LIB


def _pts(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4
    return a


# ------------------------------------------------------------------ linear algebra probes
def jacobi_svd3(A):
    A = np.ascontiguousarray(A, np.float64)
    U, s, V = np.empty((3, 3)), np.empty(3), np.empty((3, 3))
    lib().orc_jacobi_svd3(A.ctypes.data_as(C.c_void_p), U.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p))
    return U, s, V


def eigh3(A):
    A = np.ascontiguousarray(A, np.float64)
    w, V = np.empty(3), np.empty((3, 3))
    lib().orc_eigh3(A.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p))
    return w, V


def inverse4(A):
    A = np.ascontiguousarray(A, np.float64)
    out = np.empty((4, 4))
    lib().orc_inverse4(A.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def ldlt6_solve(A, b):
    A = np.ascontiguousarray(A, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    x = np.empty(6)
    lib().orc_ldlt6_solve(A.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p))
    return x


def so3_exp(omega):
    w = np.ascontiguousarray(omega, np.float64)
    R = np.empty((3, 3))
    lib().orc_so3_exp(w.ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p))
    return R


# ------------------------------------------------------------------ kNN / covariances
def knn(points, queries, k, brute=False, num_threads=0):
    p, q = _pts(points), _pts(queries)
    idx = np.empty((q.shape[0], k), np.int32)
    d2 = np.empty((q.shape[0], k), np.float32)
    if brute:
        lib().orc_knn_bruteforce(p, p.shape[0], q, q.shape[0], k, idx, d2)
    else:
        lib().orc_knn(p, p.shape[0], q, q.shape[0], k, idx, d2, num_threads)
    return idx, d2


def covariances_from_knn(points, knn_idx, method=REG_PLANE):
    p = _pts(points)
    idx = np.ascontiguousarray(knn_idx, np.int32)
    covs = np.empty((p.shape[0], 4, 4), np.float64)
    lib().orc_covariances_from_knn(p, p.shape[0], idx, idx.shape[1], method, covs.reshape(-1))
    return covs


class FastGICP:
    """Mirror of fast_gicp::FastGICP driven through the CPU oracle (row-major matrices)."""

    def __init__(self, max_iterations=64, rotation_epsilon=2e-3, transformation_epsilon=5e-4, corr_dist=np.finfo(np.float32).max,
                 k=20, regularization=REG_PLANE, optimizer=OPT_LM, lm_max_iterations=10, lm_init_lambda_factor=1e-9, num_threads=0):
        self._h = C.c_void_p(lib().orc_gicp_create())
        self.params = dict(max_iterations=max_iterations, rotation_epsilon=rotation_epsilon, transformation_epsilon=transformation_epsilon,
                           corr_dist=corr_dist, k=k, regularization=regularization, optimizer=optimizer,
                           lm_max_iterations=lm_max_iterations, lm_init_lambda_factor=lm_init_lambda_factor, num_threads=num_threads)
        self._apply()
        self.n_src = self.n_tgt = 0

    def _apply(self):
        p = self.params
        lib().orc_gicp_set_params(self._h, p["max_iterations"], p["rotation_epsilon"], p["transformation_epsilon"], p["corr_dist"], p["k"],
                                  p["regularization"], p["optimizer"], p["lm_max_iterations"], p["lm_init_lambda_factor"], p["num_threads"])

    def set(self, **kw):
        self.params.update(kw)
        self._apply()

    def __del__(self):
        try:
            lib().orc_gicp_destroy(self._h)
        except Exception:
            pass

    def setInputSource(self, pts):
        p = _pts(pts)
        self.n_src = p.shape[0]
        lib().orc_gicp_set_source(self._h, p, p.shape[0])

    def setInputTarget(self, pts):
        p = _pts(pts)
        self.n_tgt = p.shape[0]
        lib().orc_gicp_set_target(self._h, p, p.shape[0])

    def setSourceCovariances(self, covs):
        c = np.ascontiguousarray(covs, np.float64).reshape(-1)
        lib().orc_gicp_set_source_covs(self._h, c, c.size // 16)

    def setTargetCovariances(self, covs):
        c = np.ascontiguousarray(covs, np.float64).reshape(-1)
        lib().orc_gicp_set_target_covs(self._h, c, c.size // 16)

    def getSourceCovariances(self):
        lib().orc_gicp_ensure_covariances(self._h)
        c = np.empty((self.n_src, 4, 4))
        lib().orc_gicp_get_source_covs(self._h, c.reshape(-1))
        return c

    def getTargetCovariances(self):
        lib().orc_gicp_ensure_covariances(self._h)
        c = np.empty((self.n_tgt, 4, 4))
        lib().orc_gicp_get_target_covs(self._h, c.reshape(-1))
        return c

    def linearize(self, T, want_Hb=True):
        T = np.ascontiguousarray(T, np.float64).reshape(-1)
        if want_Hb:
            H, b = np.empty((6, 6)), np.empty(6)
            e = lib().orc_gicp_linearize(self._h, T, H.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
            return e, H, b
        return lib().orc_gicp_linearize(self._h, T, None, None)

    def compute_error(self, T):
        return lib().orc_gicp_compute_error(self._h, np.ascontiguousarray(T, np.float64).reshape(-1))

    def correspondences(self):
        corr = np.empty(self.n_src, np.int32)
        d2 = np.empty(self.n_src, np.float32)
        lib().orc_gicp_get_correspondences(self._h, corr, d2)
        return corr, d2

    def align(self, guess=None, want_points=False):
        g = np.eye(4, dtype=np.float32) if guess is None else np.ascontiguousarray(guess, np.float32)
        T = np.empty((4, 4), np.float32)
        res = np.zeros(4, np.int32)
        H = np.empty(36)
        out = np.empty((self.n_src, 4), np.float32) if want_points else None
        secs = lib().orc_gicp_align(self._h, g.reshape(-1), T.reshape(-1), out.ctypes.data_as(C.c_void_p) if want_points else None, res, H)
        self.last = dict(converged=bool(res[0]), iterations=int(res[1]), n_linearize=int(res[2]), n_compute_error=int(res[3]),
                         seconds=secs, final_hessian=H.reshape(6, 6), points=out)
        return T

    def getFitnessScore(self, max_range=np.finfo(np.float64).max):
        return lib().orc_gicp_fitness(self._h, max_range)


DIRECT27, DIRECT7, DIRECT1 = 0, 1, 2
ADDITIVE, ADDITIVE_WEIGHTED, MULTIPLICATIVE = 0, 1, 2


class FastVGICP(FastGICP):
    """Mirror of fast_gicp::FastVGICP (voxelised GICP) driven through the CPU oracle."""

    def __init__(self, resolution=1.0, search_method=DIRECT1, voxel_mode=ADDITIVE, **kw):
        L = lib()
        L.orc_vgicp_create.restype = C.c_void_p
        L.orc_vgicp_destroy.argtypes = [C.c_void_p]
        L.orc_vgicp_set_voxel_params.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int]
        L.orc_vgicp_set_target.argtypes = [C.c_void_p, _f32p, C.c_int]
        L.orc_vgicp_linearize.argtypes = [C.c_void_p, _f64p, C.c_void_p, C.c_void_p]
        L.orc_vgicp_linearize.restype = C.c_double
        L.orc_vgicp_compute_error.argtypes = [C.c_void_p, _f64p]
        L.orc_vgicp_compute_error.restype = C.c_double
        L.orc_vgicp_num_correspondences.argtypes = [C.c_void_p]
        L.orc_vgicp_get_voxels.argtypes = [C.c_void_p, C.c_int, _i32p, _i32p, _f64p, _f64p]
        L.orc_vgicp_align.argtypes = [C.c_void_p, _f32p, _f32p, _i32p]
        L.orc_vgicp_align.restype = C.c_double
        self._h = C.c_void_p(L.orc_vgicp_create())
        self.params = dict(max_iterations=64, rotation_epsilon=2e-3, transformation_epsilon=5e-4, corr_dist=np.finfo(np.float32).max,
                           k=20, regularization=REG_PLANE, optimizer=OPT_LM, lm_max_iterations=10, lm_init_lambda_factor=1e-9, num_threads=0)
        self.params.update(kw)
        self._apply()
        self.n_src = self.n_tgt = 0
        L.orc_vgicp_set_voxel_params(self._h, resolution, search_method, voxel_mode)

    def __del__(self):
        try:
            lib().orc_vgicp_destroy(self._h)
        except Exception:
            pass

    def setInputTarget(self, pts):
        p = _pts(pts)
        self.n_tgt = p.shape[0]
        lib().orc_vgicp_set_target(self._h, p, p.shape[0])

    def linearize(self, T, want_Hb=True):
        T = np.ascontiguousarray(T, np.float64).reshape(-1)
        if want_Hb:
            H, b = np.empty((6, 6)), np.empty(6)
            e = lib().orc_vgicp_linearize(self._h, T, H.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
            return e, H, b
        return lib().orc_vgicp_linearize(self._h, T, None, None)

    def compute_error(self, T):
        return lib().orc_vgicp_compute_error(self._h, np.ascontiguousarray(T, np.float64).reshape(-1))

    def num_correspondences(self):
        return lib().orc_vgicp_num_correspondences(self._h)

    def voxels(self):
        cap = max(self.n_tgt, 1)
        coords, num = np.zeros((cap, 3), np.int32), np.zeros(cap, np.int32)
        mean, cov = np.zeros((cap, 3)), np.zeros((cap, 6))
        nv = lib().orc_vgicp_get_voxels(self._h, cap, coords.reshape(-1), num, mean.reshape(-1), cov.reshape(-1))
        return coords[:nv], num[:nv], mean[:nv], cov[:nv]

    def align(self, guess=None):
        g = np.eye(4, dtype=np.float32) if guess is None else np.ascontiguousarray(guess, np.float32)
        T = np.empty((4, 4), np.float32)
        res = np.zeros(4, np.int32)
        secs = lib().orc_vgicp_align(self._h, g.reshape(-1), T.reshape(-1), res)
        self.last = dict(converged=bool(res[0]), iterations=int(res[1]), n_linearize=int(res[2]), n_compute_error=int(res[3]), seconds=secs)
        return T


# ------------------------------------------------------------------ A-LOAM features
class _FeatArrays(C.Structure):
    _fields_ = ([(n, C.c_void_p) for n in (
        "cloud", "src_index", "intensity_num", "range_vec", "scan_angle", "curvature", "inten_curvature", "curvature2",
        "distance_source", "other_source", "neighbor_picked", "inten_neighbor_picked", "label", "inten_label", "ground_marked",
        "corner_sharp", "corner_less_sharp", "surf_flat", "surf_less_flat", "inten_sharp", "inten_less_sharp", "ground_points",
        "corner_sharp_w", "surf_flat_w", "inten_sharp_w", "scan_start", "scan_end")]
        + [("counts", C.c_int * 16), ("groundparam", C.c_double * 11), ("ground_evals", C.c_double * 3)])


_FEAT_F32 = {"range_vec", "scan_angle", "curvature", "inten_curvature", "curvature2", "distance_source", "other_source",
             "corner_sharp_w", "surf_flat_w", "inten_sharp_w"}


def extract_features(scan_xyzi, n_scans=16, min_range=0.5, max_range=80.0, use_intensity=1):
    """Returns a dict of numpy arrays mirroring the reference's per-point arrays and feature clouds."""
    x = _pts(scan_xyzi)
    n = x.shape[0]
    cap = n + 8
    arrs = {}
    fa = _FeatArrays()
    for name, _ in _FeatArrays._fields_[:27]:
        if name == "cloud":
            a = np.zeros((cap, 4), np.float32)
        elif name in ("scan_start", "scan_end"):
            a = np.zeros(64, np.int32)
        elif name == "ground_points":
            a = np.zeros(cap * 10, np.int32)  # duplicates allowed (scanRegistration.cpp:333-348)
        elif name in _FEAT_F32:
            a = np.zeros(cap, np.float32)
        else:
            a = np.zeros(cap, np.int32)
        arrs[name] = a
        setattr(fa, name, a.ctypes.data_as(C.c_void_p))
    L = lib()
    L.orc_extract_features.argtypes = [_f32p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(_FeatArrays)]
    secs = L.orc_extract_features(x, n, n_scans, min_range, max_range, use_intensity, C.byref(fa))
    cnt = list(fa.counts)
    m = cnt[0]
    out = {"seconds": secs, "cloud_size": m, "ground_size": cnt[8], "inten_merged": cnt[9],
           "groundparam": np.array(list(fa.groundparam)), "ground_evals": np.array(list(fa.ground_evals)),
           "scan_start": arrs["scan_start"][:n_scans].copy(), "scan_end": arrs["scan_end"][:n_scans].copy()}
    per_point = ["src_index", "intensity_num", "range_vec", "scan_angle", "curvature", "inten_curvature", "curvature2", "distance_source",
                 "other_source", "neighbor_picked", "inten_neighbor_picked", "label", "inten_label", "ground_marked"]
    out["cloud"] = arrs["cloud"][:m].copy()
    for k in per_point:
        out[k] = arrs[k][:m].copy()
    for k, ci in (("corner_sharp", 1), ("corner_less_sharp", 2), ("surf_flat", 3), ("surf_less_flat", 4), ("inten_sharp", 5),
                  ("inten_less_sharp", 6), ("ground_points", 7)):
        out[k] = arrs[k][:cnt[ci]].copy()
    out["corner_sharp_w"] = arrs["corner_sharp_w"][:cnt[1]].copy()
    out["surf_flat_w"] = arrs["surf_flat_w"][:cnt[3]].copy()
    out["inten_sharp_w"] = arrs["inten_sharp_w"][:cnt[5]].copy()
    return out


def deskew(xyzi, q_wxyz, t, scan_period=0.1):
    """RGC_odometer.cpp:1441-1481 (adjustDistortion) for one cloud; xyzi: [n, 4] float32"""
    P = np.ascontiguousarray(xyzi, np.float32)
    out = np.empty_like(P)
    q = np.ascontiguousarray(q_wxyz, np.float64)
    tt = np.ascontiguousarray(t, np.float64)
    L = lib()
    L.orc_deskew.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
    L.orc_deskew.restype = None
    L.orc_deskew(P.ctypes.data, len(P), q.ctypes.data, tt.ctypes.data, scan_period, out.ctypes.data)
    return out


def voxel_grid(xyzi, leaf):
    """pcl::VoxelGrid<PointXYZI>::filter, downsample_all_data, min_points_per_voxel 0; returns [m, 4]"""
    P = np.ascontiguousarray(xyzi, np.float32)
    out = np.empty_like(P)
    pt = C.c_int(0)
    L = lib()
    L.orc_voxel_grid.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    L.orc_voxel_grid.restype = C.c_int
    m = L.orc_voxel_grid(P.ctypes.data, len(P), leaf, out.ctypes.data, C.byref(pt))
    return out[:m].copy()


def _assoc(fn_name, map_pts, feats, q_wxyz, t, n_out2):
    M, F = _pts(map_pts), np.ascontiguousarray(feats, np.float32)
    q, tt = np.ascontiguousarray(q_wxyz, np.float64), np.ascontiguousarray(t, np.float64)
    n = len(F)
    valid = np.zeros(n, np.int32)
    o1 = np.zeros((n, 3), np.float64)
    o2 = np.zeros((n, 3) if n_out2 == 3 else (n,), np.float64)
    fn = getattr(lib(), fn_name)
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    fn.restype = None
    fn(M.ctypes.data, len(M), F.ctypes.data, n, q.ctypes.data, tt.ctypes.data, valid.ctypes.data, o1.ctypes.data, o2.ctypes.data)
    return valid.astype(bool), o1, o2


def assoc_edges(map_pts, feats, q_wxyz, t):
    """RGC_mapping.cpp:1093-1136: (valid, point_a, point_b) per edge feature (x, y, z, weight)"""
    return _assoc("orc_assoc_edges", map_pts, feats, q_wxyz, t, 3)


def assoc_planes(map_pts, feats, q_wxyz, t):
    """RGC_mapping.cpp:1192-1240: (valid, unit normal, negative_OA_dot_norm) per planar feature"""
    return _assoc("orc_assoc_planes", map_pts, feats, q_wxyz, t, 1)


def colpiv_qr_solve_5x3(A, b):
    A, b = np.ascontiguousarray(A, np.float64), np.ascontiguousarray(b, np.float64)
    x = np.zeros(3)
    L = lib()
    L.orc_colpiv_qr_solve_5x3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_colpiv_qr_solve_5x3.restype = None
    L.orc_colpiv_qr_solve_5x3(A.ctypes.data, b.ctypes.data, x.ctypes.data)
    return x


def max_threads() -> int:
    return lib().orc_max_threads()
