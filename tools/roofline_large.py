"""HBM-roofline evidence for the covariance / linearize / compute_error kernels at sizes that do not
fit in L2 (the >= 60 % bar is only meaningful for the batched / large-map configs).
Target = the C2 submap replicated on a grid of offsets (n_tiles x 500k points); source = the sweep
replicated the same way, so every source point has a correspondence.
    python tools/roofline_large.py [n_tiles=16]  -> one JSON line
Under ncu:  ncu --set full -k regex:"k_covariance|k_linearize|k_compute_error" ... python tools/roofline_large.py 16"""
import json
import os
import sys

os.environ.setdefault("RGC_NO_OVERLAP", "1")  # clean per-stage times: source and target on one stream

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import rgc_slam_b200 as rgc

n_tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 16
pairs = bench.build_workload(0, bench.N_SUBMAP, 1)
tgt0, src0 = pairs[0]["tgt"], pairs[0]["src"]
side = int(np.ceil(np.sqrt(n_tiles)))
offs = [(400.0 * (i % side), 400.0 * (i // side)) for i in range(n_tiles)]
tgt = np.concatenate([tgt0 + np.array([ox, oy, 0, 0], np.float32) for ox, oy in offs], 0)
# source: a big, slightly perturbed copy of part of the target (every point has a correspondence),
# so the accumulate kernels stream hundreds of MB
n_src_tiles = max(1, n_tiles // 4)
src = tgt[: n_src_tiles * len(tgt0)].copy()
src[:, :3] += np.random.default_rng(0).normal(0, 0.01, (len(src), 3)).astype(np.float32)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists("MEASURED_PEAKS.json") else 6650.0
ctx = rgc.Context(0)
ext = torch.cuda.ExternalStream(ctx.stream)
g = rgc.FastGICP(ctx)
g.setMaxCorrespondenceDistance(2.0)
g.setGridCell(0.1)   # 18-bit grid limit: 400 m tiles x 4 need a coarser finest voxel
g.setTargetCovarianceMode(False)  # all target covariances: k_covariance over 8 M points is one of the kernels measured here
dt = torch.from_numpy(tgt).cuda()
ds = torch.from_numpy(src).cuda()
g.setInputTarget(dt)
g.setInputSource(ds)
e, H, b = g.linearize(np.eye(4))  # builds covariances (timed by the library's events)
st = g.stage_ms()
n_t, n_s, k = len(tgt), len(src), 20


def timed(fn, reps=5):
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        fn()
        e1.record(ext)
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


ctx.set_profiling(True)
T = np.eye(4)
lin_total = timed(lambda: g.linearize(T))
km = []
for _ in range(5):
    g.linearize(T)
    g.compute_error(T)
    km.append(ctx.last_kernel_ms())
corr_ms = float(np.median([x["k_correspond"] for x in km]))
lin_ms = float(np.median([x["k_linearize"] for x in km]))
ce_ms = float(np.median([x["k_compute_error"] for x in km]))
out = {"n_target": n_t, "n_source": n_s, "peak_GBps": peak,
       "k_knn_tile": {"ms": st["tgt_knn"], "Mqueries_per_s": n_t / st["tgt_knn"] / 1e3},
       "k_covariance": {"ms": st["tgt_cov"], "alg_bytes": n_t * (16 + 4 * k + 48), "GBps": n_t * (16 + 4 * k + 48) / st["tgt_cov"] / 1e6},
       "k_correspond": {"ms": corr_ms, "Mqueries_per_s": n_s / corr_ms / 1e3},
       "k_linearize": {"ms": lin_ms, "alg_bytes": n_s * 184, "GBps": n_s * 184 / lin_ms / 1e6},
       "k_compute_error": {"ms": ce_ms, "alg_bytes": n_s * 84, "GBps": n_s * 84 / ce_ms / 1e6},
       "tgt_build_ms": st["tgt_build"], "inliers": int((g.correspondences()[0] >= 0).sum())}
for kk in ("k_covariance", "k_linearize", "k_compute_error"):
    out[kk]["frac_of_peak"] = out[kk]["GBps"] / peak
print(json.dumps(out))
