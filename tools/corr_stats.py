import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import rgc_slam_b200 as rgc
pairs = bench.build_workload(0, bench.N_SUBMAP, 1)
ctx = rgc.Context(0)
g = bench.new_reg(rgc, ctx)
g.setInputTarget(pairs[0]["tgt"]); g.setInputSource(pairs[0]["src"])
T = np.ascontiguousarray(pairs[0]["guess"].astype(np.float64).T)
L = rgc.lib(); L.rgc_debug_correspond_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
n = len(pairs[0]["src"]); st = np.zeros((n, 4), np.int64)
for rep in range(3):
    ctx.check(L.rgc_debug_correspond_stats(g._h, T.ctypes.data, st.ctypes.data))
cyc, nodes, look, cand = st.T
for nm, v in (("cycles", cyc), ("nodes", nodes), ("lookups", look), ("cands", cand)):
    print(f"{nm:8s} mean {v.mean():9.1f} p50 {np.percentile(v,50):8.0f} p90 {np.percentile(v,90):8.0f} p99 {np.percentile(v,99):8.0f} p99.9 {np.percentile(v,99.9):8.0f} max {v.max():8d}")
o = np.argsort(cyc)[-6:]
print("slowest:", [(int(cyc[i]), int(nodes[i]), int(look[i]), int(cand[i])) for i in o])
w = cyc.reshape(-1)[: (n // 32) * 32].reshape(-1, 32)
print("per-warp(32 consecutive) max cycles: mean %.0f p50 %.0f p99 %.0f max %d" % (w.max(1).mean(), np.percentile(w.max(1), 50), np.percentile(w.max(1), 99), w.max()))
print("cycles per lookup+cand (mean)", cyc.mean() / (look.mean() + cand.mean()))
# who is slow: distance of the found neighbour, and whether there is one at all (after 3 reps the search is hinted)
corr, d2 = g.correspondences()
inv = np.empty(n, np.int64)
# stats are in SORTED source order; correspondences come back in original order: compare distributions only
print("matched %.3f, d2 quantiles (matched) p50 %.4f p90 %.4f p99 %.4f max %.3f" % ((corr >= 0).mean(), *np.percentile(d2[corr >= 0], [50, 90, 99]), d2[corr >= 0].max()))
print("cycle quantiles p99.9/p50 = %.1f ; share of total cycles in the slowest 1%% of queries: %.3f" % (np.percentile(cyc, 99.9) / np.percentile(cyc, 50), np.sort(cyc)[-n // 100:].sum() / cyc.sum()))
