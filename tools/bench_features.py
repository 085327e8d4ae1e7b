"""Config C3: A-LOAM curvature + edge/planar feature extraction + ground-plane fit over batches of
synthetic 16/32-beam scans (BASELINE.json configs[2]).  Reports scans/s for batch sizes 1/64/1024 on
the GPU (device time of all kernels, CUDA events; and wall time incl. H2D/D2H of every output array)
next to the single-threaded CPU oracle (the reference node is single-threaded).  One JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import rgc_slam_b200 as rgc
from oracle import oracle as orc
from rgc_slam_b200 import synth
from rgc_slam_b200.features import extract_features

n_distinct = int(sys.argv[1]) if len(sys.argv) > 1 else 16
scene = synth.Scene.make(synth.BASE_SEED + 3000)
traj = synth.trajectory(n_distinct + 4, seed=3)
out = {}
for beams, az in ((16, 1800), (32, 900)):
    scans = [synth.lidar_scan(scene, traj[f], n_beams=beams, n_azimuth=az, seed=synth.BASE_SEED + 3000 + f) for f in range(n_distinct)]
    pts = float(np.mean([len(s) for s in scans]))
    t0 = time.perf_counter()
    for s in scans[:8]:
        orc.extract_features(s, n_scans=beams)
    cpu = 8 / (time.perf_counter() - t0)
    res = {"points_per_scan": pts, "cpu_scans_per_s_1thread": cpu}
    for B in (1, 64, 1024):
        batch = [scans[i % n_distinct] for i in range(B)]
        extract_features(batch[: min(B, 4)], n_rings=beams)  # warm the pool
        dev, wall = [], []
        for rep in range(3):
            t0 = time.perf_counter()
            r, ms = extract_features(batch, n_rings=beams)
            wall.append(time.perf_counter() - t0)
            dev.append(ms)
        t0 = time.perf_counter()
        r, ms = extract_features(batch, n_rings=beams, want_arrays=False)
        lists_only = time.perf_counter() - t0
        res[f"batch{B}"] = {"device_ms": float(np.median(dev)), "scans_per_s_device": B / (np.median(dev) * 1e-3),
                            "scans_per_s_wall_all_outputs": B / float(np.median(wall)), "scans_per_s_wall_lists_only": B / lists_only,
                            "pass_b_GBps": B * pts * 44 / (np.median(dev) * 1e-3) / 1e9}
    out[f"{beams}-beam"] = res
print(json.dumps(out))
