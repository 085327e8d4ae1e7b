"""TEST-ONLY: builds and loads tests/hostsim/libhostsim.so (g++ build of the product's host/device
headers).  Not part of the product; see hostsim.cpp."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "libhostsim.so")
        src = os.path.join(HERE, "hostsim.cpp")
        csrc = os.path.join(HERE, "..", "..", "rgc_slam_b200", "csrc")
        deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".hpp"))]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            subprocess.check_call([cxx, "-O2", "-g", "-fopenmp", "-fPIC", "-std=c++17", "-ffp-contract=off", "-shared", "-o", so, src])
        L = C.CDLL(so)
        L.sim_knn.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_int, _i32p, _f32p, C.c_float, C.c_void_p]
        L.sim_covs.argtypes = [_f32p, C.c_int, _i32p, C.c_int, C.c_int, _f64p]
        L.sim_linearize.argtypes = [_f32p, C.c_int, _f32p, C.c_int, _f64p, _f64p, _f64p, C.c_float, C.c_float, _f64p, _f64p, _f64p, _i32p]
        L.sim_solve_ldlt6.argtypes = [_f64p, _f64p, _f64p]
        L.sim_se3_delta.argtypes = [_f64p, _f64p]
        L.sim_is_converged.argtypes = [_f64p, C.c_double, C.c_double]
        L.sim_heap64_topk.argtypes = [np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS"), C.c_int, C.c_int, C.c_int,
                                      np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")]
        L.sim_deskew.argtypes = [_f32p, C.c_int, _f64p, _f64p, C.c_float, _f32p]
        L.sim_voxel_grid.argtypes = [_f32p, C.c_int, C.c_float, _f32p]
        L.sim_voxel_grid.restype = C.c_int
        L.sim_map_assoc.argtypes = [_f32p, C.c_int, _f32p, C.c_int, _f64p, _f64p, C.c_int, _i32p, _f64p, _f64p]
        _LIB = L
    return _LIB


def knn(points, queries, k, cell=0.0, want_stats=False):
    p = np.ascontiguousarray(points, np.float32)
    q = np.ascontiguousarray(queries, np.float32)
    idx = np.empty((len(q), k), np.int32)
    d2 = np.empty((len(q), k), np.float32)
    stats = (C.c_longlong * 3)()
    lib().sim_knn(p, len(p), q, len(q), k, idx, d2, cell, stats)
    if want_stats:
        return idx, d2, [s / max(len(q), 1) for s in stats]
    return idx, d2


def covs(points, knn_idx, method):
    p = np.ascontiguousarray(points, np.float32)
    i = np.ascontiguousarray(knn_idx, np.int32)
    out = np.empty((len(p), 4, 4))
    lib().sim_covs(p, len(p), i, i.shape[1], method, out.reshape(-1))
    return out


def linearize(src, tgt, cov_a, cov_b, T, thr, cell=0.0):
    s = np.ascontiguousarray(src, np.float32)
    t = np.ascontiguousarray(tgt, np.float32)
    e, H, b = np.zeros(1), np.zeros(36), np.zeros(6)
    corr = np.zeros(len(s), np.int32)
    lib().sim_linearize(s, len(s), t, len(t), np.ascontiguousarray(cov_a, np.float64).reshape(-1), np.ascontiguousarray(cov_b, np.float64).reshape(-1),
                        np.ascontiguousarray(T, np.float64).reshape(-1), thr, cell, e, H, b, corr)
    return e[0], H.reshape(6, 6), b, corr


def deskew(xyzi, q_wxyz, t, scan_period=0.1):
    P = np.ascontiguousarray(xyzi, np.float32)
    out = np.empty_like(P)
    lib().sim_deskew(P, len(P), np.ascontiguousarray(q_wxyz, np.float64), np.ascontiguousarray(t, np.float64), scan_period, out)
    return out


def voxel_grid(xyzi, leaf):
    P = np.ascontiguousarray(xyzi, np.float32)
    out = np.empty_like(P)
    m = lib().sim_voxel_grid(P, len(P), leaf, out)
    return out[:m].copy()


def map_assoc(map_xyz1, feats, q_wxyz, t, plane):
    M = np.ascontiguousarray(map_xyz1, np.float32)
    F = np.ascontiguousarray(feats, np.float32)
    n = len(F)
    valid = np.zeros(n, np.int32)
    o1 = np.zeros((n, 3), np.float64)
    o2 = np.zeros(n if plane else (n, 3), np.float64)
    lib().sim_map_assoc(M, len(M), F, n, np.ascontiguousarray(q_wxyz, np.float64), np.ascontiguousarray(t, np.float64), int(bool(plane)), valid, o1, o2)
    return valid.astype(bool), o1, o2


def heap64_topk(keys, k, stride=1):
    K = np.ascontiguousarray(keys, np.uint64)
    out = np.zeros(k, np.uint64)
    lib().sim_heap64_topk(K, len(K), k, stride, out)
    return out
