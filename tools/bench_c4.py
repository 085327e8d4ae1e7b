"""Config C4: independent loop-closure style registrations, one VLP-16 sweep vs a
100 000-point submap each, initial error U(+-0.5 m, +-5 deg) about the truth.  Pairs are independent
units: they are sharded over ranks (torchrun) with no collective, and inside a rank over T host
threads, each with its own rgc context (stream pair + memory pool) so that the latency-bound LM
loops of different pairs overlap on the GPU.

    python tools/bench_c4.py [--pairs 256] [--threads 4]            # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_c4.py --pairs 4096

Prints one JSON line: pairs/s over all ranks (max-over-ranks wall time), success rate against the
known truth, per-thread-count sweep when --sweep is given."""
import argparse
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import rgc_slam_b200 as rgc
from rgc_slam_b200 import batch, sharded, workloads

N_SUBMAP_C4 = workloads.N_SUBMAP_C4


def recovered(results, pairs):
    ok = 0
    for T, p in zip(results, pairs):
        E = np.linalg.inv(p["truth"]) @ np.asarray(T, np.float64)
        ang = np.arccos(np.clip((np.trace(E[:3, :3]) - 1) / 2, -1, 1))
        ok += int(np.linalg.norm(E[:3, 3]) < 0.05 and ang < np.deg2rad(0.5))
    return ok


def run_batched(pairs, device, chunk=0, pinned=True):
    """one rgc_batch_align call over all pairs (include/rgc_batch.h); host clouds (pinned by default), H2D inside"""
    import torch
    ctx = rgc.Context(device)
    prm = batch.default_params()
    prm.max_iterations, prm.max_correspondence_distance = 64, 2.0
    if pinned:
        cache = {}
        def pin(a):
            if id(a) not in cache:
                cache[id(a)] = torch.from_numpy(a).pin_memory()
            return cache[id(a)]
        pp = [dict(src=pin(p["src"]), tgt=pin(p["tgt"]), guess=p["guess"]) for p in pairs]
    else:
        pp = pairs
    batch.align_batch(pp, ctx=ctx, params=prm)  # warm-up = the same call once: the pools' large blocks (two chunks in flight) are cudaMalloc'ed here
    ctx.synchronize()
    t0 = time.perf_counter()
    res = batch.align_batch(pp, ctx=ctx, params=prm, want_fitness=True, max_chunk_pairs=chunk)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    stages = batch.last_stage_ms(ctx)
    ok = recovered([r["T"] for r in res], pairs)
    accepted = sum(int(r["converged"] and r["fitness"] <= 0.1) for r in res)  # RGC_mapping.cpp:2070-2071
    ctx.close()
    return dt, ok, accepted, stages


def run(pairs, n_threads, device):
    results = [None] * len(pairs)
    ctxs = [rgc.Context(device) for _ in range(n_threads)]

    def worker(tid):
        ctx = ctxs[tid]
        for i in range(tid, len(pairs), n_threads):
            p = pairs[i]
            g = rgc.FastGICP(ctx)  # a fresh object per registration, like the reference's callers
            g.setMaximumIterations(64)
            g.setMaxCorrespondenceDistance(2.0)
            g.setInputTarget(p["tgt"])
            g.setInputSource(p["src"])
            T = g.align(p["guess"])
            results[i] = (np.asarray(T, np.float64), g.hasConverged())
            g = None

    # warm-up: pools, module load
    for tid in range(n_threads):
        p = pairs[tid % len(pairs)]
        g = rgc.FastGICP(ctxs[tid])
        g.setMaxCorrespondenceDistance(2.0)
        g.setInputTarget(p["tgt"])
        g.setInputSource(p["src"])
        g.align(p["guess"])
        g = None
    for c in ctxs:
        c.synchronize()
    th = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    for c in ctxs:
        c.synchronize()
    dt = time.perf_counter() - t0
    ok = recovered([T for T, _ in results], pairs)
    for c in ctxs:
        c.close()
    return dt, ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=256, help="total pairs over all ranks")
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--sweep", action="store_true", help="also time 1, 2, 4, 8 threads (rank 0, single GPU)")
    ap.add_argument("--batched", action="store_true", help="one rgc_batch_align call per rank instead of the host-thread farm")
    ap.add_argument("--chunk", type=int, default=0, help="pairs per chunk of the batched path (0 = library default)")
    ap.add_argument("--pageable", action="store_true", help="batched: pass pageable numpy clouds instead of pinned tensors")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    lo, hi = sharded.shard_range(args.pairs, world, rank)
    pairs = workloads.make_c4_pairs(lo, hi - lo)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    stages = accepted = None
    if args.batched:
        dt, ok, accepted, stages = run_batched(pairs, local, args.chunk, not args.pageable)
    else:
        dt, ok = run(pairs, args.threads, local)
    if world > 1:
        import torch
        t = torch.tensor([dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        o = torch.tensor([ok], device="cuda")
        dist.all_reduce(o, op=dist.ReduceOp.SUM)
        dt, ok = float(t.item()), int(o.item())
    out = {"config": "C4", "pairs": args.pairs, "n_gpus": world, "threads_per_gpu": args.threads, "n_source": int(len(pairs[0]["src"])),
           "n_target": N_SUBMAP_C4, "seconds": dt, "pairs_per_s": args.pairs / dt, "ms_per_pair_per_gpu": 1e3 * dt / (hi - lo),
           "recovered_truth": ok, "scaling": "strong" if world > 1 else None,
           "mode": "batched (rgc_batch_align)" if args.batched else "host-thread farm", "accepted_rank0": accepted, "stage_ms_rank0": stages,
           "inputs": ("host (pageable numpy)" if (args.pageable or not args.batched) else "host (pinned)") + ", H2D inside the timed region"}
    if args.sweep and world == 1:
        out["thread_sweep"] = {}
        for nt in (1, 2, 4, 8):
            d, _ = run(pairs, nt, local)
            out["thread_sweep"][str(nt)] = len(pairs) / d
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
