"""Pre-step timing on the C2 shapes: raw VLP-16 sweep (de-skew + 0.2 m voxel filter) and
raw 500k-point submap (0.3 m voxel filter) in front of one registration.
  fused   raw host clouds -> setInput*Filtered (de-skew / filter / sort / kNN / cov on the device) -> align
  staged  rgc.deskew / rgc.voxel_grid back to the host, then setInput* -> align
  cpu     the oracle's de-skew + two voxel filters (single thread, like pcl::VoxelGrid)
Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
import rgc_slam_b200 as rgc
from oracle import oracle as orc

p = bench.build_workload(0, bench.N_SUBMAP, 1)[0]
rng = np.random.default_rng(0)


def stamp(P):
    Q = P.copy()
    Q[:, 3] = rng.integers(0, 16, len(Q)) + np.float32(0.1) * rng.uniform(0, 1, len(Q)).astype(np.float32)
    return Q


S, T = stamp(p["src"]), stamp(p["tgt"])
q = np.array([0.99995, 0.001, -0.002, 0.009])
q /= np.linalg.norm(q)
t = np.array([0.12, 0.01, -0.004])
ctx = rgc.Context(0)


def xyz1(V):
    W = V.copy()
    W[:, 3] = 1.0
    return W


def fused():
    g = bench.new_reg(rgc, ctx)
    nt = g.setInputTargetFiltered(T[:], 0.3)
    ns = g.setInputSourceFiltered(S[:], 0.2, q, t)
    g.align(p["guess"])
    return ns, nt, g.last_result["iterations"]


def staged():
    Tg = rgc.voxel_grid(T, 0.3, ctx)
    Sg = rgc.voxel_grid(rgc.deskew(S, q, t, ctx=ctx), 0.2, ctx)
    g = bench.new_reg(rgc, ctx)
    g.setInputTarget(xyz1(Tg))
    g.setInputSource(xyz1(Sg))
    g.align(p["guess"])


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ctx.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ctx.synchronize()
        ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))


ns, nt, it = fused()
out = {"n_source_raw": len(S), "n_target_raw": len(T), "n_source_filtered": ns, "n_target_filtered": nt, "lm_iterations": it,
       "fused_ms": timed(fused), "staged_ms": timed(staged),
       "gpu_voxel_grid_500k_ms": timed(lambda: rgc.voxel_grid(T, 0.3, ctx)), "gpu_deskew_sweep_ms": timed(lambda: rgc.deskew(S, q, t, ctx=ctx))}
t0 = time.perf_counter()
orc.voxel_grid(T, 0.3)
out["cpu_voxel_grid_500k_ms"] = 1e3 * (time.perf_counter() - t0)
t0 = time.perf_counter()
orc.voxel_grid(orc.deskew(S, q, t), 0.2)
out["cpu_deskew_plus_voxel_grid_sweep_ms"] = 1e3 * (time.perf_counter() - t0)
print(json.dumps(out))
