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
                                           "inten_less_sharp", "n_inten_less_sharp",
                                           "surf_less_flat", "n_surf_less_flat", "ground_points")]
                + [("ground_cap", C.c_int32), ("device_ms", C.c_float)])


def extract_features(scans, n_rings=16, min_range=0.5, max_range=80.0, use_intensity=1, ctx: api.Context | None = None,
                     want_arrays=True, ground_cap=None):
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
    # the two unbounded clouds: surfPointsLessFlatScan and GroundPoints (a sample is appended up to 10 times)
    mk("surf_less_flat", total_out, np.int32); mk("n_surf_less_flat", nb, np.int32)
    gcap = int(ground_cap) if ground_cap is not None else 10 * max(len(s) for s in scans)
    mk("ground_points", (nb, gcap), np.int32)
    out.ground_cap = gcap
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
        # surfPointsLessFlatScan (:586-592) and GroundPoints (:338, duplicates in push order)
        d["surf_less_flat"] = arrs["surf_less_flat"][o0:o0 + int(arrs["n_surf_less_flat"][b])]
        d["ground_points"] = arrs["ground_points"][b, :min(d["ground_size"], gcap)]
        res.append(d)
    return res, float(out.device_ms)


class FeatureExtractor:
    """Batched feature extraction with PREALLOCATED PINNED buffers (config C3 end to end): the raw scans are
    concatenated once into a pinned input buffer, the outputs a front end consumes — the five feature lists with
    their weights, the per-point labels, cloud sizes / ring ranges and the ground parameters — land in pinned
    output buffers; nothing is allocated per call and no pageable copy is made.  `extract_features` above
    returns every intermediate array and is what the parity tests use."""

    def __init__(self, ctx: api.Context | None = None, n_rings=16, max_scans=1024, max_points=1024 * 30000, min_range=0.5, max_range=80.0,
                 use_intensity=1, labels=True):
        import torch
        self.ctx = ctx or api.default_context(0)
        self.n_rings, self.max_scans, self.max_points = n_rings, max_scans, max_points
        self.prm = (min_range, max_range, use_intensity)
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()  # noqa: E731
        i32, f32, f64 = torch.int32, torch.float32, torch.float64
        self.raw = pin((max_points, 4), f32)
        self.offs = pin((max_scans + 1,), i32)
        self.buf = {"cloud_size": pin((max_scans,), i32), "scan_start": pin((max_scans, 64), i32), "scan_end": pin((max_scans, 64), i32),
                    "groundparam": pin((max_scans, 11), f64), "ground_size": pin((max_scans,), i32), "inten_merged": pin((max_scans,), i32)}
        if labels:
            self.buf["label"] = pin((max_points + 8 * max_scans,), i32)
        for name, per_seg, has_w in _LISTS:
            self.buf[name] = pin((max_scans, n_rings * 6 * per_seg), i32)
            self.buf["n_" + name] = pin((max_scans,), i32)
            if has_w:
                self.buf[name + "_w"] = pin((max_scans, n_rings * 6 * per_seg), f32)
        self.n = 0
        self.total = 0
        L = api.lib()
        L.rgc_feat_extract.argtypes = [C.c_void_p, C.POINTER(_Batch), C.POINTER(_Out)]

    def load(self, scans):
        """concatenate `scans` (list of (n_i, 4) float32 arrays) into the pinned input buffer"""
        import torch
        self.n = len(scans)
        assert self.n <= self.max_scans
        o = 0
        offs = [0]
        for s in scans:
            m = len(s)
            assert o + m <= self.max_points
            self.raw[o:o + m] = torch.from_numpy(np.ascontiguousarray(s[:, :4], np.float32))
            o += m
            offs.append(o)
        self.offs[:self.n + 1] = torch.tensor(offs, dtype=torch.int32)
        self.total = o

    @property
    def h2d_bytes(self):
        return 16 * self.total + 4 * (self.n + 1)

    @property
    def d2h_bytes(self):
        nb, tot = self.n, self.total + 8 * self.n
        b = 0
        for k, t in self.buf.items():
            per = t[0].numel() * t.element_size() if t.dim() > 1 else t.element_size()
            b += (tot if k == "label" else nb) * per
        return b

    def run(self) -> float:
        """one rgc_feat_extract call on the loaded batch; returns the device time (ms) of its kernels"""
        batch = _Batch(self.raw.data_ptr(), self.offs.data_ptr(), self.n, self.n_rings, self.prm[0], self.prm[1], self.prm[2])
        out = _Out()
        for k, t in self.buf.items():
            setattr(out, k, t.data_ptr())
        self.ctx.check(api.lib().rgc_feat_extract(self.ctx._h, C.byref(batch), C.byref(out)))
        return float(out.device_ms)

    def result(self, b: int) -> dict:
        """views of scan b's outputs (valid until the next run)"""
        o0 = int(self.offs[b]) + 8 * b
        m = int(self.buf["cloud_size"][b])
        d = {"cloud_size": m, "groundparam": self.buf["groundparam"][b].numpy(), "ground_size": int(self.buf["ground_size"][b]),
             "inten_merged": int(self.buf["inten_merged"][b]), "scan_start": self.buf["scan_start"][b, :self.n_rings].numpy(),
             "scan_end": self.buf["scan_end"][b, :self.n_rings].numpy()}
        if "label" in self.buf:
            d["label"] = self.buf["label"][o0:o0 + m].numpy()
        for name, per_seg, has_w in _LISTS:
            c = int(self.buf["n_" + name][b])
            d[name] = self.buf[name][b, :c].numpy()
            if has_w:
                d[name + "_w"] = self.buf[name + "_w"][b, :c].numpy()
        return d

    def close(self):
        self.buf = {}
        self.raw = self.offs = None
