"""Config C5 (BASELINE.json configs[4]): one 128-beam sweep (~230k points) registered against a
very large global map that is slab-sharded across the ranks, with one NCCL all-reduce of the 29
partial doubles per linearize.  Run under torchrun; rank 0 prints one JSON line.
    torchrun --nproc-per-node 8 tools/bench_c5.py [n_tiles=100]   (n_tiles x 500k points)
The map is the C2 submap replicated on a grid of 400 m offsets; the sweep is placed in a central tile.
RGC_C5_CHECK=1 also runs the unsharded registration on rank 0's GPU and compares."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
import rgc_slam_b200 as rgc
from rgc_slam_b200 import sharded, synth


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    pairs = bench.build_workload(0, bench.N_SUBMAP, 1)
    tgt0 = pairs[0]["tgt"]
    side = int(np.ceil(np.sqrt(n_tiles)))
    offs = [(400.0 * (i % side), 400.0 * (i // side)) for i in range(n_tiles)]
    tgt = np.concatenate([tgt0 + np.array([ox, oy, 0, 0], np.float32) for ox, oy in offs], 0)
    centre = offs[(side // 2) * side + side // 2] if (side // 2) * side + side // 2 < n_tiles else offs[0]
    # 128-beam sweep from the C2 scene / pose of the pair, moved into the central tile
    scene = synth.Scene.make(synth.BASE_SEED + 2000)
    need = int(np.ceil(bench.N_SUBMAP * 1.03 / 20000.0)) + 2
    traj = synth.trajectory(need + bench.N_PAIRS + 8, seed=2)
    src = synth.to_xyz1(synth.lidar_scan(scene, traj[need], n_beams=128, seed=4242))
    src[:, 0] += centre[0]
    src[:, 1] += centre[1]
    guess = pairs[0]["guess"].astype(np.float64).copy()
    # conjugate the guess by the tile offset so it is the same relative motion about the moved sweep
    Toff = np.eye(4)
    Toff[:2, 3] = centre
    guess = (Toff @ guess @ np.linalg.inv(Toff)).astype(np.float32)
    ctx = rgc.Context(local)
    out = {"world": world, "n_target": int(len(tgt)), "n_source": int(len(src))}

    def params(g):
        g.setMaximumIterations(25)
        g.setMaxCorrespondenceDistance(2.0)
        g.setTransformationEpsilon(1e-6)
        g.setGridCell(0.1)

    gs = sharded.ShardedFastGICP(ctx, cov_halo=4.0)
    params(gs)
    t0 = time.perf_counter()
    gs.setInputTarget(tgt)
    gs.setInputSource(src)
    ctx.synchronize()
    t_set = time.perf_counter() - t0
    t0 = time.perf_counter()
    T = gs.align(guess)
    ctx.synchronize()
    t_align_cold = time.perf_counter() - t0     # includes the local kNN + covariances of the slab
    t0 = time.perf_counter()
    T2 = gs.align(guess)
    ctx.synchronize()
    t_align_warm = time.perf_counter() - t0     # target structures cached: LM iterations + all-reduces only
    lin = []
    Tg = guess.astype(np.float64)
    for _ in range(5):
        t0 = time.perf_counter()
        e, H, b = gs.linearize(Tg)
        lin.append(time.perf_counter() - t0)
    # all-reduce latency alone (29 doubles on the library's stream)
    buf = torch.zeros(32, dtype=torch.float64, device=f"cuda:{local}")
    ar = []
    if world > 1:
        for _ in range(20):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dist.all_reduce(buf[:29])
            torch.cuda.synchronize()
            ar.append(time.perf_counter() - t0)
    stats = torch.tensor([t_set, t_align_cold, t_align_warm, float(np.median(lin)), float(len(gs.local_target))], dtype=torch.float64, device=f"cuda:{local}")
    mx = stats.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    out.update(set_target_s_max=mx[0].item(), align_cold_s_max=mx[1].item(), align_warm_s_max=mx[2].item(), linearize_ms_max=mx[3].item() * 1e3,
               local_target_max=int(mx[4].item()), iterations=gs.last_result["iterations"], converged=gs.hasConverged(), n_allreduce=gs.n_allreduce,
               allreduce_us_median=float(np.median(ar) * 1e6) if ar else None, stage_ms_rank0=gs.stage_ms(), inliers=gs.last_inliers())
    if rank == 0 and os.environ.get("RGC_C5_CHECK"):
        gu = rgc.FastGICP(ctx)
        params(gu)
        gu.setInputTarget(tgt)
        gu.setInputSource(src)
        t0 = time.perf_counter()
        Tu = gu.align(guess)
        ctx.synchronize()
        eu, Hu, bu = gu.linearize(Tg)
        out.update(unsharded_align_cold_s=time.perf_counter() - t0, unsharded_stage_ms=gu.stage_ms(),
                   H_rel=float(np.abs(H - Hu).max() / np.abs(Hu).max()), b_rel=float(np.abs(b - bu).max() / np.abs(bu).max()),
                   pose_dt=float(np.abs(T[:3, 3] - Tu[:3, 3]).max()), pose_dR=float(np.abs(T[:3, :3] - Tu[:3, :3]).max()),
                   iters_equal=gs.last_result["iterations"] == gu.last_result["iterations"])
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
