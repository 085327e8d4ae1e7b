"""Python mirror of the A-LOAM feature path (include/rgc_features.h): batched
ScanRegistration::laserCloudHandler numerics (rgc_slam/src/scanRegistration.cpp:110-663) on the GPU.
Output keys follow the reference's array names so the parity tests read like its code."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api

_PER_POINT_F32 = ["range_vec", "scan_angle", "curvature", "inten_curvature", "curvature2", "distance_source", "other_source"]
_PER_POINT_I32 = ["src_index", "intensity_num", "label", "inten_label", "neighbor_picked", "inten_neighbor_picked", "ground_marked"]
_LISTS = [("corner_sharp", 20, True), ("corner_less_sharp", 22, False), ("surf_flat", 40, True), ("inten_sharp", 20, True), ("inten_less_sharp", 21, False)]


class _Batch(C.Structure):
    _fields_ = [("xyzi", C.c_void_p), ("scan_offsets", C.c_void_p), ("n_scans", C.c_int), ("n_rings", C.c_int),
                ("minimum_range", C.c_double), ("maximum_range", C.c_double), ("use_intensity", C.c_int)]


class _Out(C.Structure):
    _fields_ = ([(n, C.c_void_p) for n in ("cloud_size", "scan_start", "scan_end", "groundparam", "ground_size", "inten_merged", "cloud",
                                           "src_index", "intensity_num", "range_vec", "scan_angle", "curvature", "inten_curvature", "curvature2",
                                           "distance_source", "other_source", "label", "inten_label", "neighbor_picked", "inten_neighbor_picked",
                                           "ground_marked",
                                           "corner_sharp", "corner_sharp_w", "n_corner_sharp",
                                           "corner_less_sharp", "n_corner_less_sharp",
                                           "surf_flat", "surf_flat_w", "n_surf_flat",
                                           "inten_sharp", "inten_sharp_w", "n_inten_sharp",
                                           "inten_less_sharp", "n_inten_less_sharp")]
                + [("device_ms", C.c_float)])


def extract_features(scans, n_rings=16, min_range=0.5, max_range=80.0, use_intensity=1, ctx: api.Context | None = None,
                     want_arrays=True):
    """scans: list of (n_i, 4) float32 arrays (x, y, z, intensity) in firing order.
    Returns (list of per-scan dicts, device_ms)."""
    ctx = ctx or api.default_context(0)
    L = api.lib()
    L.rgc_feat_extract.argtypes = [C.c_void_p, C.POINTER(_Batch), C.POINTER(_Out)]
    nb = len(scans)
    offs = np.zeros(nb + 1, np.int32)
    offs[1:] = np.cumsum([len(s) for s in scans])
    raw = np.ascontiguousarray(np.concatenate([np.asarray(s, np.float32)[:, :4] for s in scans], 0), np.float32)
    total_out = int(offs[-1]) + 8 * nb
    batch = _Batch(raw.ctypes.data, offs.ctypes.data, nb, n_rings, min_range, max_range, use_intensity)
    out = _Out()
    arrs = {}

    def mk(name, shape, dtype):
        a = np.zeros(shape, dtype)
        arrs[name] = a
        setattr(out, name, a.ctypes.data)

    mk("cloud_size", nb, np.int32); mk("scan_start", (nb, 64), np.int32); mk("scan_end", (nb, 64), np.int32)
    mk("groundparam", (nb, 11), np.float64); mk("ground_size", nb, np.int32); mk("inten_merged", nb, np.int32)
    if want_arrays:
        mk("cloud", (total_out, 4), np.float32)
        for n in _PER_POINT_F32:
            mk(n, total_out, np.float32)
        for n in _PER_POINT_I32:
            mk(n, total_out, np.int32)
    for name, per_seg, has_w in _LISTS:
        mk(name, (nb, n_rings * 6 * per_seg), np.int32)
        mk("n_" + name, nb, np.int32)
        if has_w:
            mk(name + "_w", (nb, n_rings * 6 * per_seg), np.float32)
    ctx.check(L.rgc_feat_extract(ctx._h, C.byref(batch), C.byref(out)))
    res = []
    for b in range(nb):
        m = int(arrs["cloud_size"][b])
        o0 = int(offs[b]) + 8 * b
        d = {"cloud_size": m, "ground_size": int(arrs["ground_size"][b]), "inten_merged": int(arrs["inten_merged"][b]),
             "groundparam": arrs["groundparam"][b].copy(), "scan_start": arrs["scan_start"][b, :n_rings].copy(),
             "scan_end": arrs["scan_end"][b, :n_rings].copy()}
        if want_arrays:
            d["cloud"] = arrs["cloud"][o0:o0 + m]
            for n in _PER_POINT_F32 + _PER_POINT_I32:
                d[n] = arrs[n][o0:o0 + m]
        for name, per_seg, has_w in _LISTS:
            c = int(arrs["n_" + name][b])
            d[name] = arrs[name][b, :c]
            if has_w:
                d[name + "_w"] = arrs[name + "_w"][b, :c]
        # surfPointsLessFlatScan (:586-592): every point of a processed sextant whose label is <= 0
        res.append(d)
    return res, float(out.device_ms)
