"""How many target covariances does each LM iteration of a cold C2 align request on demand?
Runs the align with max_iterations = 1, 2, ... on fresh objects and counts the claimed target points.
    python tools/need_stats.py [pairs=3]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
import rgc_slam_b200 as rgc

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
pairs = bench.build_workload(0, bench.N_SUBMAP, n_pairs)
ctx = rgc.Context(0)
for pi, p in enumerate(pairs):
    prev, row = 0, []
    for it in range(1, 9):
        g = bench.new_reg(rgc, ctx)
        g.setMaximumIterations(it)
        g.setInputTarget(p["tgt"])
        g.setInputSource(p["src"])
        g.align(p["guess"])
        _, st = g.target_cov_state()
        claimed = int((st != 0).sum())
        row.append((it, g.last_result["iterations"], g.last_result["n_linearize"], claimed, claimed - prev))
        prev = claimed
        if g.hasConverged():
            break
    print(f"pair {pi}: n_src {len(p['src'])}  (max_it, iterations, n_linearize, claimed total, new) =", row)
