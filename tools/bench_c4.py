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

import bench
import rgc_slam_b200 as rgc
from rgc_slam_b200 import sharded, synth

N_SUBMAP_C4 = 100_000


def perturbation(rng):
    """U(+-0.5 m, +-5 deg) about the identity"""
    t = rng.uniform(-0.5, 0.5, 3)
    w = np.deg2rad(rng.uniform(-5.0, 5.0, 3))
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    R = np.eye(3) + (np.sin(th) / th) * K + ((1 - np.cos(th)) / th**2) * K @ K if th > 1e-12 else np.eye(3)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def make_pairs(rank, n_pairs, n_base=4):
    base = bench.build_workload(rank, N_SUBMAP_C4, n_base)
    rng = np.random.Generator(np.random.PCG64(synth.BASE_SEED + 4000 + rank))
    out = []
    for i in range(n_pairs):
        b = base[i % n_base]
        guess = (perturbation(rng) @ b["truth"]).astype(np.float32)
        out.append(dict(src=b["src"], tgt=b["tgt"], guess=guess, truth=b["truth"]))
    return out


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
    ok = 0
    for (T, conv), p in zip(results, pairs):
        E = np.linalg.inv(p["truth"]) @ T
        ang = np.arccos(np.clip((np.trace(E[:3, :3]) - 1) / 2, -1, 1))
        ok += int(np.linalg.norm(E[:3, 3]) < 0.05 and ang < np.deg2rad(0.5))
    for c in ctxs:
        c.close()
    return dt, ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=256, help="total pairs over all ranks")
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--sweep", action="store_true", help="also time 1, 2, 4, 8 threads (rank 0, single GPU)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    lo, hi = sharded.shard_range(args.pairs, world, rank)
    pairs = make_pairs(rank, hi - lo)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
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
           "recovered_truth": ok, "scaling": "weak" if world > 1 else None, "inputs": "host (pageable numpy), H2D inside the timed region"}
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
