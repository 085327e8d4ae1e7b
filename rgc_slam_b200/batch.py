"""ctypes binding of include/rgc_batch.h: batched independent registrations (config C4, loop-closure
candidate verification — the caller's loop is rgc_slam/src/RGC_mapping.cpp:2051-2086, one
registration + `hasConverged() && getFitnessScore() <= 0.1` gate per candidate)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api


class _Pair(C.Structure):
    _fields_ = [("source", C.c_void_p), ("n_source", C.c_size_t), ("source_stride", C.c_size_t),
                ("target", C.c_void_p), ("n_target", C.c_size_t), ("target_stride", C.c_size_t), ("guess", C.c_float * 16)]


class _PairResult(C.Structure):
    _fields_ = [("final_T", C.c_float * 16), ("result", api._Result), ("fitness", C.c_double)]


def _lib():
    L = api.lib()
    L.rgc_batch_align.argtypes = [C.c_void_p, C.POINTER(api._Params), C.POINTER(_Pair), C.c_size_t, C.c_int, C.c_double, C.c_int, C.POINTER(_PairResult)]
    L.rgc_batch_last_stage_ms.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    return L


def default_params() -> api._Params:
    p = api._Params()
    api.lib().rgc_params_default(C.byref(p))
    return p


def align_batch(pairs, ctx: api.Context | None = None, params: api._Params | None = None, want_fitness: bool = True,
                fitness_max_range: float = float(np.finfo(np.float64).max), max_chunk_pairs: int = 0):
    """pairs: sequence of dicts / tuples (src, tgt, guess) with host clouds [n, >=3] float32 (numpy arrays or
    pinned / pageable torch CPU tensors) and a row-major 4x4 guess (None = identity).
    Returns a list of dicts {T (4x4 float32 row-major), converged, iterations, n_linearize, n_compute_error,
    n_inliers, final_error, final_hessian (6x6), fitness}."""
    ctx = ctx or api.default_context(0)
    n = len(pairs)
    arr = (_Pair * n)()
    keep = []
    for i, p in enumerate(pairs):
        src, tgt, guess = (p["src"], p["tgt"], p.get("guess")) if isinstance(p, dict) else p
        sp, sn, ss, sdev, sk = api._as_cloud(src)
        tp, tn, ts, tdev, tk = api._as_cloud(tgt)
        if sdev or tdev:
            raise ValueError("align_batch takes host clouds")
        keep += [sk, tk]
        g = np.eye(4, dtype=np.float32) if guess is None else np.asarray(guess, np.float32)
        arr[i].source, arr[i].n_source, arr[i].source_stride = sp, sn, ss
        arr[i].target, arr[i].n_target, arr[i].target_stride = tp, tn, ts
        arr[i].guess[:] = np.ascontiguousarray(g.T).reshape(-1).tolist()  # column-major
    out = (_PairResult * n)()
    prm = params if params is not None else default_params()
    ctx.check(_lib().rgc_batch_align(ctx._h, C.byref(prm), arr, n, int(bool(want_fitness)), float(fitness_max_range), int(max_chunk_pairs), out))
    res = []
    for i in range(n):
        r = out[i].result
        res.append(dict(T=np.array(out[i].final_T, np.float32).reshape(4, 4).T.copy(), converged=bool(r.converged), iterations=r.iterations,
                        n_linearize=r.n_linearize, n_compute_error=r.n_compute_error, n_inliers=r.n_inliers, final_error=r.final_error,
                        final_hessian=np.array(r.final_hessian).reshape(6, 6).T.copy(), fitness=out[i].fitness, device_ms=r.device_ms))
    return res


def last_stage_ms(ctx: api.Context | None = None):
    ctx = ctx or api.default_context(0)
    ms = np.zeros(6, np.float32)
    rounds = C.c_int(0)
    _lib().rgc_batch_last_stage_ms(ctx._h, ms.ctypes.data, C.byref(rounds))
    d = dict(zip(("upload_ingest", "source_build", "target_build", "source_knn_cov", "lm_rounds", "fitness"), ms.tolist()))
    d["rounds"] = rounds.value
    return d
