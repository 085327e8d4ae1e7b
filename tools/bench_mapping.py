"""Mapping-node association timing: one frame = build the corner and surface map search
structures, then 2 solver iterations x (current + last frame) x (edge + planar) association calls
(RGC_mapping.cpp:1073-1290).  GPU wall clock (host arrays in / out) vs the oracle on all host threads.
Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

import rgc_slam_b200 as rgc
from oracle import oracle as orc
from test_oracle_mapping import _pose, _scene

corner, surf, rng = _scene(3)
corner = np.tile(corner, (4, 1))                          # ~10k corner points
corner[:, :3] += np.repeat(rng.uniform(-60, 60, (4, 3)) * [1, 1, 0], len(corner) // 4, 0).astype(np.float32)
surf = np.tile(surf, (5, 1))                              # ~90k surface points
surf[:, :3] += np.repeat(rng.uniform(-60, 60, (5, 3)) * [1, 1, 0], len(surf) // 5, 0).astype(np.float32)
rot, q, t = _pose(rng)
fe = np.zeros((1500, 4), np.float32)
fe[:, :3] = rot.inv().apply(corner[rng.choice(len(corner), len(fe)), :3] + rng.normal(0, 0.05, (len(fe), 3)) - t)
fp = np.zeros((8000, 4), np.float32)
fp[:, :3] = rot.inv().apply(surf[rng.choice(len(surf), len(fp)), :3] + rng.normal(0, 0.03, (len(fp), 3)) - t)
ctx = rgc.Context(0)


def gpu_frame():
    mc, ms = rgc.FeatureMap(corner, ctx), rgc.FeatureMap(surf, ctx)
    nv = 0
    for _ in range(2):
        for _ in range(2):
            nv += mc.associate_edges(fe, q, t)[0].sum()
            nv += ms.associate_planes(fp, q, t)[0].sum()
    mc.close()
    ms.close()
    return nv


def cpu_frame():
    nv = 0
    for _ in range(2):
        for _ in range(2):
            nv += orc.assoc_edges(corner, fe, q, t)[0].sum()     # the oracle rebuilds its kd-tree per call,
            nv += orc.assoc_planes(surf, fp, q, t)[0].sum()      # the reference once per frame: see cpu_ms_tree_once
    return nv


def timed(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))


nv = gpu_frame()
out = {"n_corner_map": len(corner), "n_surf_map": len(surf), "n_edge_features": len(fe), "n_planar_features": len(fp), "valid_per_frame": int(nv),
       "gpu_frame_ms": timed(gpu_frame, 20), "cpu_frame_ms_incl_8_tree_builds": timed(cpu_frame, 3), "cpu_threads": orc.max_threads()}
m = rgc.FeatureMap(surf, ctx)
out["gpu_planes_call_ms"] = timed(lambda: m.associate_planes(fp, q, t), 20)
out["gpu_map_build_ms"] = timed(lambda: rgc.FeatureMap(surf, ctx).close(), 20)
print(json.dumps(out))
