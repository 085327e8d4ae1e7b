"""Per-warp statistics of the tile kNN kernel on the C2 submap (debug aid)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import rgc_slam_b200 as rgc
pairs = bench.build_workload(0, bench.N_SUBMAP, 1)
ctx = rgc.Context(0)
L = rgc.lib()
L.rgc_debug_tile_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.c_float]
for name in ("tgt", "src"):
    P = pairs[0][name]
    nw = (len(P) + 31) // 32
    st = np.zeros((nw, 4), np.int64)
    for rep in range(2):
        ctx.check(L.rgc_debug_tile_stats(ctx._h, P.ctypes.data, len(P), 16, 20, st.ctypes.data, 0.0))
    cyc, nodes, cands, ins = st.T.copy()
    lca = nodes >> 40
    nodes = nodes & ((1 << 40) - 1)
    start = (ins >> 32) * 64e-3  # us
    start = start - start.min()
    ins = ins & 0xffffffff
    os.makedirs("gpurun_out", exist_ok=True)
    np.savez_compressed(f"gpurun_out/tile_stats_{name}.npz", cyc=cyc, nodes=nodes, cands=cands, ins=ins, lca=lca, start=start)
    print(name, "warps", nw)
    for nm, v in (("cycles", cyc), ("nodes", nodes), ("cands", cands), ("fold steps", ins)):
        print(f"  {nm:10s} mean {v.mean():10.1f} p50 {np.percentile(v,50):10.0f} p90 {np.percentile(v,90):10.0f} p99 {np.percentile(v,99):10.0f} max {v.max():10d} sum {v.sum():14d}")
    o = np.argsort(cyc)[-5:]
    print("  slowest warps:", [(int(i), int(cyc[i]), int(nodes[i]), int(cands[i]), int(ins[i])) for i in o])
    print("  corr(cycles, cands) %.2f  corr(cycles, nodes) %.2f" % (np.corrcoef(cyc, cands)[0, 1], np.corrcoef(cyc, nodes)[0, 1]))
