"""Profiling driver: a few cold C2 steps (same step as bench.py) with nothing else around them,
so `ncu` launch lists / full captures stay short.  Usage: python tools/prof_step.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import rgc_slam_b200 as rgc

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
eager = len(sys.argv) > 2 and sys.argv[2] == "eager"  # all target covariances at the first align (the reference's schedule)
pairs = bench.build_workload(0, bench.N_SUBMAP, 1)
ctx = rgc.Context(0)
dev = [dict(src=torch.from_numpy(p["src"]).cuda(), tgt=torch.from_numpy(p["tgt"]).cuda()) for p in pairs]
torch.cuda.synchronize()
for i in range(steps):
    g = bench.new_reg(rgc, ctx, on_demand=not eager)
    g.setInputTarget(dev[0]["tgt"][:])
    g.setInputSource(dev[0]["src"][:])
    T = g.align(pairs[0]["guess"])
    print(i, g.last_result, g.stage_ms())
    fs = g.getFitnessScore()
ctx.synchronize()
