"""Multi-GPU check (run under torchrun through `gpurun --gpus N`, not collected by pytest):
  C5: one registration against a slab-sharded target (selected on the device from the full cloud), ONE
      ncclAllReduce per LM step issued by the library — final pose / H / b must equal (a) the single-GPU
      unsharded result and (b) the CPU ORACLE;
  C4: independent pairs sharded across ranks through rgc_batch_align — results identical to a single-rank run.
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
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rgc_slam_b200 as rgc  # noqa: E402
from rgc_slam_b200 import batch, sharded, synth, workloads  # noqa: E402


def rot_angle(Ra, Rb):
    R = Ra.astype(np.float64).T @ Rb.astype(np.float64)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
    return float(np.arctan2(np.linalg.norm(v), (np.trace(R) - 1) / 2))


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
    # ---------------- C5 sharded, library-owned NCCL communicator
    gs = sharded.ShardedFastGICP(ctx, cov_halo=4.0)
    params(gs)
    t0 = time.perf_counter()
    gs.setInputTarget(tgt, want_index=True)
    gs.setInputSource(src)
    gs.waitInputs()
    ctx.synchronize()
    out["set_inputs_s"] = time.perf_counter() - t0
    Tl = np.eye(4)
    Tl[:3, 3] = [0.2, 0.05, 0.0]
    es, Hs, bs = gs.linearize(Tl)
    t0 = time.perf_counter()
    Ts = gs.align(guess)
    ctx.synchronize()
    t_sharded = time.perf_counter() - t0
    n_ar0 = gs.n_allreduce
    Ts2 = gs.align(guess)
    fit_s = gs.getFitnessScore(4.0)
    out.update(sharded_align_s=t_sharded, sharded_iters=gs.last_result["iterations"], allreduces_per_align=gs.n_allreduce - n_ar0 - 1,
               n_linearize=gs.last_result["n_linearize"], n_compute_error=gs.last_result["n_compute_error"],
               local_target=gs.n_local, allreduce=gs.allreduce_kind, allreduce_us=gs.allreduce_us(), repeat_identical=bool(np.array_equal(Ts, Ts2)))
    # the device-side slab selection keeps exactly the points the host-side rule keeps, in input order
    axis, lo, hi = gs.slab
    keep = sharded.slab_select(tgt, axis, lo, hi, 2.0 + 4.0)
    out["slab_selection_ok"] = bool(np.array_equal(keep.astype(np.int32), gs.local_index))
    # identical on every rank?
    Tt = torch.from_numpy(Ts.astype(np.float64)).cuda()
    Tmax, Tmin = Tt.clone(), Tt.clone()
    dist.all_reduce(Tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(Tmin, op=dist.ReduceOp.MIN)
    out["ranks_agree"] = bool((Tmax == Tmin).all().item())
    # ---------------- the same with the host-callback transport (torch.distributed): same numbers
    gc = sharded.ShardedFastGICP(ctx, cov_halo=4.0, use_nccl=False)
    params(gc)
    gc.setInputTarget(tgt)
    gc.setInputSource(src)
    Tc = gc.align(guess)
    out["callback_transport_pose_equal"] = bool(np.abs(Tc - Ts).max() < 1e-6)
    out["callback_allreduces_per_align"] = gc.n_allreduce
    gc.close()
    gc = None
    # ---------------- unsharded on rank 0's GPU, and the CPU oracle
    if rank == 0:
        gu = rgc.FastGICP(ctx)
        params(gu)
        gu.setInputTarget(tgt)
        gu.setInputSource(src)
        eu, Hu, bu = gu.linearize(Tl)
        t0 = time.perf_counter()
        Tu = gu.align(guess)
        ctx.synchronize()
        out.update(unsharded_align_s=time.perf_counter() - t0,
                   lin_err_rel=abs(es - eu) / abs(eu), lin_H_rel=float(np.abs(Hs - Hu).max() / np.abs(Hu).max()),
                   lin_b_rel=float(np.abs(bs - bu).max() / np.abs(bu).max()),
                   pose_dt=float(np.abs(Ts[:3, 3] - Tu[:3, 3]).max()), pose_dR=float(np.abs(Ts[:3, :3] - Tu[:3, :3]).max()),
                   iters_equal=gs.last_result["iterations"] == gu.last_result["iterations"],
                   fitness_rel=abs(fit_s - gu.getFitnessScore(4.0)) / gu.getFitnessScore(4.0))
        if os.environ.get("RGC_SKIP_ORACLE") is None:
            from oracle import oracle as orc
            o = orc.FastGICP(max_iterations=25, corr_dist=2.0, transformation_epsilon=1e-6)
            o.setInputTarget(tgt)
            o.setInputSource(src)
            eo, Ho, bo = o.linearize(Tl)
            To = o.align(guess)
            out.update(oracle_lin_H_rel=float(np.abs(Hs - Ho).max() / np.abs(Ho).max()), oracle_lin_b_rel=float(np.abs(bs - bo).max() / np.abs(bo).max()),
                       oracle_pose_dt=float(np.abs(Ts[:3, 3] - To[:3, 3]).max()), oracle_pose_angle=rot_angle(Ts[:3, :3], To[:3, :3]),
                       oracle_iters_equal=gs.last_result["iterations"] == o.last["iterations"], oracle_converged_equal=gs.hasConverged() == o.last["converged"])
    gs.close()
    # ---------------- C4: independent pairs through the batch path
    n_pairs = 16
    lo_p, hi_p = sharded.shard_range(n_pairs, world, rank)
    pairs = workloads.make_c4_pairs(lo_p, hi_p - lo_p)
    prm = batch.default_params()
    prm.max_iterations, prm.max_correspondence_distance = 64, 2.0
    mine = [(p["index"], r["T"].tolist(), r["converged"], r["fitness"]) for p, r in zip(pairs, batch.align_batch(pairs, ctx=ctx, params=prm))]
    allres = sharded.gather_results(mine, world)
    if rank == 0:
        ok = [r[0] for r in allres] == list(range(n_pairs))
        full = batch.align_batch(workloads.make_c4_pairs(0, n_pairs), ctx=ctx, params=prm)
        for (i, T, conv, fit), r in zip(allres, full):
            ok = ok and np.array_equal(np.asarray(T, np.float32), r["T"]) and conv == r["converged"] and fit == r["fitness"]
        out["c4_identical_to_single_rank"] = bool(ok)
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
