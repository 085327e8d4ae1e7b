"""Multi-GPU check (run under torchrun through `gpurun --gpus N`, not collected by pytest):
  C5: one registration against a slab-sharded target, NCCL all-reduce of the partial H/b per
      linearize — final pose / H / b must equal the single-GPU unsharded result;
  C4: independent pairs sharded across ranks — results identical to a single-rank run.
Prints one JSON line from rank 0."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rgc_slam_b200 as rgc  # noqa: E402
from rgc_slam_b200 import sharded, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_map = int(os.environ.get("RGC_MAP_POINTS", "2000000"))
    scene = synth.Scene.make(synth.BASE_SEED + 5000)
    traj = synth.trajectory(140, seed=5)
    # a larger map: many sweeps accumulated in the frame of pose 100 (same seeds on every rank)
    chunks, f = [], 100
    total = 0
    while total < n_map and f >= 0:
        sc = synth.lidar_scan(scene, traj[f], seed=9000 + f)
        Tr = synth.relative_pose(traj[f], traj[100])
        chunks.append((sc[:, :3].astype(np.float64) @ Tr[:3, :3].T + Tr[:3, 3]).astype(np.float32))
        total += len(sc)
        f -= 1
    tgt = np.ones((total, 4), np.float32)
    tgt[:, :3] = np.concatenate(chunks, 0)
    src = synth.to_xyz1(synth.lidar_scan(scene, traj[101], n_beams=64 if n_map > 1_000_000 else 16, seed=777))
    guess = np.eye(4, dtype=np.float32)
    ctx = rgc.Context(local)

    def params(g):
        g.setMaximumIterations(25)
        g.setMaxCorrespondenceDistance(2.0)
        g.setTransformationEpsilon(1e-6)

    out = {"world": world, "n_target": int(total), "n_source": int(len(src))}
    # ---------------- C5 sharded
    gs = sharded.ShardedFastGICP(ctx, cov_halo=4.0)
    params(gs)
    gs.setInputTarget(tgt)
    gs.setInputSource(src)
    Tl = np.eye(4)
    Tl[:3, 3] = [0.2, 0.05, 0.0]
    es, Hs, bs = gs.linearize(Tl)
    t0 = time.perf_counter()
    Ts = gs.align(guess)
    ctx.synchronize()
    t_sharded = time.perf_counter() - t0
    fit_s = gs.getFitnessScore(4.0)
    out.update(sharded_align_s=t_sharded, sharded_iters=gs.last_result["iterations"], n_allreduce=gs.n_allreduce,
               local_target=int(len(gs.local_target)), sharded_stage_ms=gs.stage_ms())
    # identical on every rank?
    Tt = torch.from_numpy(Ts.astype(np.float64)).cuda()
    Tmax, Tmin = Tt.clone(), Tt.clone()
    dist.all_reduce(Tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(Tmin, op=dist.ReduceOp.MIN)
    out["ranks_agree"] = bool((Tmax == Tmin).all().item())
    # ---------------- unsharded reference on rank 0's GPU
    if rank == 0:
        gu = rgc.FastGICP(ctx)
        params(gu)
        gu.setInputTarget(tgt)
        gu.setInputSource(src)
        eu, Hu, bu = gu.linearize(Tl)
        t0 = time.perf_counter()
        Tu = gu.align(guess)
        ctx.synchronize()
        out.update(unsharded_align_s=time.perf_counter() - t0, unsharded_stage_ms=gu.stage_ms(),
                   lin_err_rel=abs(es - eu) / abs(eu), lin_H_rel=float(np.abs(Hs - Hu).max() / np.abs(Hu).max()),
                   lin_b_rel=float(np.abs(bs - bu).max() / np.abs(bu).max()),
                   pose_dt=float(np.abs(Ts[:3, 3] - Tu[:3, 3]).max()), pose_dR=float(np.abs(Ts[:3, :3] - Tu[:3, :3]).max()),
                   iters_equal=gs.last_result["iterations"] == gu.last_result["iterations"],
                   fitness_rel=abs(fit_s - gu.getFitnessScore(4.0)) / gu.getFitnessScore(4.0))
    # ---------------- C4: independent pairs
    n_pairs = 8
    pairs = []
    for p in range(n_pairs):
        a = synth.to_xyz1(synth.lidar_scan(scene, traj[10 + 3 * p], n_azimuth=900, seed=100 + p))
        b = synth.to_xyz1(synth.lidar_scan(scene, traj[11 + 3 * p], n_azimuth=900, seed=200 + p))
        pairs.append((a, b))
    lo, hi = sharded.shard_range(n_pairs, world, rank, [len(a) + len(b) for a, b in pairs])
    mine = []
    for p in range(lo, hi):
        g = rgc.FastGICP(ctx)
        g.setInputTarget(pairs[p][0])
        g.setInputSource(pairs[p][1])
        mine.append((p, g.align().tolist(), g.hasConverged()))
    allres = sharded.gather_results(mine, world)
    if rank == 0:
        ok = [r[0] for r in allres] == list(range(n_pairs))
        for p, T, conv in allres:
            g = rgc.FastGICP(ctx)
            g.setInputTarget(pairs[p][0])
            g.setInputSource(pairs[p][1])
            ok = ok and np.array_equal(np.asarray(T, np.float32), g.align()) and conv == g.hasConverged()
        out["c4_identical_to_single_rank"] = bool(ok)
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
