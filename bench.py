#!/usr/bin/env python
"""bench.py — scan-to-map GICP aligns/s (VLP-16 sweep vs 500k-point submap), BASELINE.json config C2.

A "step" is what the odometry node does per LiDAR frame (rgc_slam/src/RGC_odometer.cpp:998-1011):
a fresh FastGICP object, the call-site parameters (25 iterations, corr 2 m, trans eps 1e-6),
setInputTarget(NEW 500k-point submap) + setInputSource(new sweep) + align(guess).  The target
changes every frame at the reference call site, so every step is COLD: the target's voxel hash and
every covariance the alignment reads are computed inside the timed region.

  value : aligns/s, inputs already resident in HBM, device time (CUDA events on the library's
          stream, per step; L2 flushed between steps outside the events); p50 / p99 beside it
  e2e   : aligns/s through the same public call with pinned HOST clouds: H2D of both clouds and
          the D2H of the result are inside the timed (wall-clock) region
  --impl reference : the CPU restatement of the reference's OpenMP FastGICP path (oracle/, the
          reference itself cannot be built here) on the same workload, all host threads

Beside the headline the line carries (N = 1): the reference's covariance schedule (`eager_target_covariances`),
the warm (cached target) and FastVGICP numbers, the HBM-bound kernels at 8 M / 2 M points
(`roofline.hbm_kernels`), config C3 (feature extraction, `c3`), config C4 (batched loop-closure
verification, `c4`).  N > 1 (torchrun): `value` = independent C2 registrations per rank (weak scaling, no
collective); `c4` = 4096 pairs split over the ranks by sharded.shard_range (strong scaling, no collective);
`c5` = one registration against a map slab-sharded over the ranks with one ncclAllReduce per LM step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rgc_slam_b200 import workloads  # noqa: E402

N_SUBMAP = workloads.N_SUBMAP_C2
CALL_SITE = workloads.CALL_SITE
N_PAIRS = 16  # distinct (sweep, submap) pairs cycled through
METRIC = "scan-to-map GICP aligns/sec (VLP-16 vs 500k-pt map)"
DTYPE = "f64 (H/b, covariances) + f32 (points, kNN distances)"


_T0 = time.perf_counter()


def log(msg: str):
    """progress on stderr (stdout carries the one JSON line)"""
    print(f"[bench +{time.perf_counter() - _T0:6.1f}s rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def build_workload(rank: int, n_submap: int, n_pairs: int = N_PAIRS):
    """config C2 stream (rgc_slam_b200/workloads.py): sweeps along a trajectory vs the accumulated submap"""
    return workloads.build_c2_pairs(rank, n_submap, n_pairs)


def config_dict(n_src, n_tgt, n_pairs, world):
    """the workload description: identical in both arms"""
    return {"workload": "C2: VLP-16 sweep vs 500k-point submap, cold target every step (new submap per frame as at "
                        "RGC_odometer.cpp:985-1009), call-site params 25 it / corr 2 m / trans_eps 1e-6, k=20 PLANE LM",
            "n_source": int(n_src), "n_target": int(n_tgt), "pairs_cycled": int(n_pairs),
            "sharding": "independent registrations per rank, no collective" if world > 1 else "single GPU",
            # (both arms print these two lines unchanged, so that the two config dicts compare equal)
            "l2": "GPU arm: L2 flushed between steps (256 MB memset); CPU arm: no device cache involved",
            "target_covariances": "GPU arm: computed on demand for the target points that become correspondences (identical values; "
                                  "`eager_target_covariances` in the GPU line = the reference's schedule); CPU arm: all target points, as the reference"}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region.  In-process NVML at 50 Hz (an
    `nvidia-smi -lms` subprocess was measurably perturbing the latency-bound LM loop: its queries
    take driver locks); falls back to nvidia-smi if pynvml is unavailable."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.sm, self.mx, self.reasons = [], [], set()
        self._stop = threading.Event()
        self.thr = None
        self.how = "nvml"
        self.interval = float(os.environ.get("RGC_CLOCK_INTERVAL", "0.02"))
        self._nv = None

    def prepare(self):
        """nvmlInit + handle lookup are slow and take driver locks: do them BEFORE the timed region."""
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu_index]) if vis and vis.split(",")[0].isdigit() else self.gpu_index
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self._nv = None
            self.how = f"nvml unavailable ({type(e).__name__}); one nvidia-smi query after the timed region"

    def _sample(self):
        nv, h = self._nv, self._h
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
        self.mx.append(self._max)
        try:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for bit, name in self.REASONS.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _loop(self):
        if self._nv is None:
            return
        while not self._stop.is_set():
            self._sample()
            self._stop.wait(self.interval)

    def start(self):
        self.thr = threading.Thread(target=self._loop, daemon=True)
        self.thr.start()

    def stop(self):
        self._stop.set()
        if self.thr:
            self.thr.join(timeout=3)
        if self._nv is None:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i", str(self.gpu_index)],
                                     capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                self.sm.append(float(out[0]))
                self.mx.append(float(out[1]))
            except Exception:
                pass
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "how": self.how}


def new_reg(rgc, ctx, on_demand=True):
    g = rgc.FastGICP(ctx)
    g.setMaximumIterations(CALL_SITE["max_iterations"])
    g.setMaxCorrespondenceDistance(CALL_SITE["corr_dist"])
    g.setTransformationEpsilon(CALL_SITE["transformation_epsilon"])
    g.setEuclideanFitnessEpsilon(1e-6)
    g.setRANSACIterations(0)
    g.setNumThreads(14)
    if not on_demand:
        g.setTargetCovarianceMode(False)
    return g


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ legs
def leg_cold(torch, rgc, ctx, ext, flush, pairs, clouds, steps, on_demand=True, wall=False):
    """`steps` cold aligns; per-step device ms (CUDA events on the library's stream) or, with wall=True,
    per-step wall seconds around the public call + synchronize"""
    times, iters, stage_acc, launches = [], [], {}, 0
    g = T = None
    for i in range(steps):
        with torch.cuda.stream(ext):
            flush.zero_()  # L2 flush between timed iterations (outside the timed region)
        ctx.synchronize()
        p, c = pairs[i % len(pairs)], clouds[i % len(pairs)]
        g = None  # the reference's object is stack-local: destroyed before the next frame's is built
        l0 = ctx.launch_count
        if wall:
            t0 = time.perf_counter()
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
        g = new_reg(rgc, ctx, on_demand)
        g.setInputTarget(c["tgt"][:])  # a new tensor view per step = a new cloud identity, as at the reference call site
        g.setInputSource(c["src"][:])
        T = g.align(p["guess"])
        if wall:
            ctx.synchronize()
            times.append(time.perf_counter() - t0)
        else:
            e1.record(ext)
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        launches += ctx.launch_count - l0
        iters.append(g.last_result["iterations"])
        for k, v in g.stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    return dict(times=np.array(times), iters=iters, stage={k: v / max(steps, 1) for k, v in stage_acc.items()}, launches=launches, last=(g, T))


def leg_roofline_large(torch, rgc, local, n_tiles=16):
    """HBM-bound kernels at sizes that do not fit in L2: target = n_tiles x 500k points (8 M),
    source = 2 M points with a correspondence each.  Kernel times are CUDA events recorded by the library
    around each launch; GB/s on SURVEY §8(d)'s algorithmic bytes and, separately, on the bytes of this
    layout (fp64 covariances, Mahalanobis matrices kept for compute_error)."""
    pair = workloads.build_c2_pairs(0, N_SUBMAP, 1)[0]
    tgt0 = pair["tgt"]
    side = int(np.ceil(np.sqrt(n_tiles)))
    offs = [(400.0 * (i % side), 400.0 * (i // side)) for i in range(n_tiles)]
    tgt = np.concatenate([tgt0 + np.array([ox, oy, 0, 0], np.float32) for ox, oy in offs], 0)
    src = tgt[: max(1, n_tiles // 4) * len(tgt0)].copy()
    src[:, :3] += np.random.default_rng(0).normal(0, 0.01, (len(src), 3)).astype(np.float32)
    ctx = rgc.Context(local)
    g = rgc.FastGICP(ctx)
    g.setMaxCorrespondenceDistance(2.0)
    g.setGridCell(0.1)
    g.setTargetCovarianceMode(False)  # all 8 M target covariances: this leg measures that kernel
    dt, ds = torch.from_numpy(tgt).cuda(local), torch.from_numpy(src).cuda(local)
    g.setInputTarget(dt)
    g.setInputSource(ds)
    T = np.eye(4)
    g.linearize(T)
    st = g.stage_ms()
    ctx.set_profiling(True)
    km = []
    for _ in range(5):
        g.linearize(T)
        g.compute_error(T)
        km.append(ctx.last_kernel_ms())
    ctx.set_profiling(False)
    n_t, n_s, k = len(tgt), len(src), 20
    peak, _ = peak_hbm()
    med = {kk: float(np.median([x[kk] for x in km])) for kk in km[0]}

    def row(ms, n, survey_b, layout_b):
        return {"ms": ms, "units": n, "survey_bytes_per_unit": survey_b, "layout_bytes_per_unit": layout_b,
                "GBps_survey_bytes": n * survey_b / ms / 1e6, "frac_survey_bytes": n * survey_b / ms / 1e6 / peak,
                "GBps_layout_bytes": n * layout_b / ms / 1e6, "frac_layout_bytes": n * layout_b / ms / 1e6 / peak}

    out = {"n_target": n_t, "n_source": n_s, "peak_GBps": peak,
           "k_covariance": row(st["tgt_cov"], n_t, 120, 16 + 4 * k + 48),
           "k_linearize": row(med["k_linearize"], n_s, 84, 184),
           "k_compute_error": row(med["k_compute_error"], n_s, 84, 84),
           "k_knn_tile": {"ms": st["tgt_knn"], "Mqueries_per_s": n_t / st["tgt_knn"] / 1e3, "bound": "sm"},
           "k_correspond": {"ms": med["k_correspond"], "Mqueries_per_s": n_s / med["k_correspond"] / 1e3, "bound": "sm"},
           "bytes": "survey = SURVEY.md §8(d) (6 x fp32 covariances, M never stored); layout = what these kernels move (DESIGN.md §3-4)"}
    g = None
    ctx.close()
    del dt, ds
    torch.cuda.empty_cache()
    return out


def run_c4(torch, dist, rgc, local, rank, world, total_pairs):
    """config C4: `total_pairs` sweep-vs-100k-submap pairs with U(+-0.5 m, +-5 deg) initial error, split over
    the ranks (sharded.shard_range) and aligned by ONE rgc_batch_align call per rank; pinned host clouds,
    H2D inside the timed region.  No collective in the data path."""
    from rgc_slam_b200 import batch, sharded
    lo, hi = sharded.shard_range(total_pairs, world, rank)
    pairs = workloads.make_c4_pairs(lo, hi - lo)
    ctx = rgc.Context(local)
    prm = batch.default_params()
    prm.max_iterations, prm.max_correspondence_distance = 64, 2.0
    cache = {}

    def pin(a):
        if id(a) not in cache:
            cache[id(a)] = torch.from_numpy(a).pin_memory()
        return cache[id(a)]

    pp = [dict(src=pin(p["src"]), tgt=pin(p["tgt"]), guess=p["guess"]) for p in pairs]
    batch.align_batch(pp, ctx=ctx, params=prm)  # warm-up = the same call once: the pools' large blocks (two chunks in flight) are cudaMalloc'ed here
    ctx.synchronize()
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
    t0 = time.perf_counter()
    res = batch.align_batch(pp, ctx=ctx, params=prm, want_fitness=True)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    stages = batch.last_stage_ms(ctx)
    ok = acc = 0
    for r, p in zip(res, pairs):
        E = np.linalg.inv(p["truth"]) @ r["T"].astype(np.float64)
        ang = np.arccos(np.clip((np.trace(E[:3, :3]) - 1) / 2, -1, 1))
        ok += int(np.linalg.norm(E[:3, 3]) < 0.05 and ang < np.deg2rad(0.5))
        acc += int(r["converged"] and r["fitness"] <= 0.1)  # the caller's gate, RGC_mapping.cpp:2070-2071
    ctx.close()
    t = torch.tensor([dt, float(ok), float(acc)], dtype=torch.float64, device=f"cuda:{local}")
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {"workload": "C4: sweep vs 100 000-point submap, initial error U(+-0.5 m, +-5 deg), 64 it / corr 2 m, rgc_batch_align",
            "pairs": total_pairs, "pairs_per_s": total_pairs / float(tmax[0].item()), "seconds_max_over_ranks": float(tmax[0].item()),
            "recovered_truth": int(t[1].item()), "accepted_by_fitness_gate": int(t[2].item()), "scaling": "strong" if world > 1 else None,
            "collectives": "none", "inputs": "host (pinned), H2D inside the timed region", "stage_ms_rank0": stages}


def run_c5(torch, dist, rgc, local, rank, world):
    """config C5: one 128-beam sweep against a large map slab-sharded over the ranks (6.25 M points per rank:
    50 M at 8 GPUs), one all-reduce of the partial (err, H, b) per LM step."""
    from rgc_slam_b200 import sharded
    n_tiles = max(2, int(round(12.5 * world)))
    case = workloads.make_c5_case(n_tiles)
    tgt_pinned = torch.from_numpy(case["tgt"]).pin_memory()
    ctx = rgc.Context(local)
    gs = sharded.ShardedFastGICP(ctx, cov_halo=4.0)
    gs.setMaximumIterations(25)
    gs.setMaxCorrespondenceDistance(2.0)
    gs.setTransformationEpsilon(1e-6)
    gs.setGridCell(0.1)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    gs.setInputTarget(tgt_pinned)
    gs.setInputSource(case["src"])
    gs.waitInputs()
    ctx.synchronize()
    t_set = time.perf_counter() - t0
    t0 = time.perf_counter()
    T = gs.align(case["guess"])
    ctx.synchronize()
    t_cold = time.perf_counter() - t0
    n0 = gs.n_allreduce
    t0 = time.perf_counter()
    T = gs.align(case["guess"])
    ctx.synchronize()
    t_warm = time.perf_counter() - t0
    n_ar = gs.n_allreduce - n0
    main_iters, main_conv = gs.last_result["iterations"], gs.hasConverged()
    Tg = case["guess"].astype(np.float64)
    lin = []
    for _ in range(5):
        t0 = time.perf_counter()
        e, H, b = gs.linearize(Tg)
        lin.append(time.perf_counter() - t0)
    ar_us = gs.allreduce_us()
    # the call-site criterion (step < 1e-6 m) is not met in 25 iterations at this size: correspondences keep flipping
    # and move the optimum by more than that; the library default (5e-4 m, lsq_registration_impl.hpp:13) is
    gs.setTransformationEpsilon(5e-4)
    gs.align(case["guess"])
    default_eps = {"iterations": gs.last_result["iterations"], "converged": gs.hasConverged()}
    gs.setTransformationEpsilon(1e-6)
    stats = torch.tensor([t_set, t_cold, t_warm, float(np.median(lin)), float(gs.n_local)], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    out = {"workload": f"C5: 128-beam sweep ({len(case['src'])} pts) vs {len(case['tgt'])}-point map slab-sharded over {world} ranks, corr 2 m, 25 it",
           "n_target": int(len(case["tgt"])), "n_source": int(len(case["src"])), "local_target_max": int(stats[4].item()),
           "set_target_s_max": float(stats[0].item()), "set_target": "full map from pinned host memory, slab + halo selected on the device, voxel hash built",
           "align_cold_ms_max": 1e3 * float(stats[1].item()), "align_warm_ms_max": 1e3 * float(stats[2].item()),
           "linearize_ms_max": 1e3 * float(stats[3].item()), "iterations": main_iters, "converged": main_conv,
           "allreduces_per_align": n_ar, "allreduce": gs.allreduce_kind, "allreduce_us": ar_us, "with_default_trans_eps_5e-4": default_eps}
    if rank == 0:  # unsharded on rank 0's GPU: H, b and the final pose must agree
        gu = rgc.FastGICP(ctx)
        gu.setMaximumIterations(25)
        gu.setMaxCorrespondenceDistance(2.0)
        gu.setTransformationEpsilon(1e-6)
        gu.setGridCell(0.1)
        gu.setInputTarget(tgt_pinned)
        gu.setInputSource(case["src"])
        t0 = time.perf_counter()
        Tu = gu.align(case["guess"])
        ctx.synchronize()
        out["unsharded_align_cold_ms"] = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter()
        Tu = gu.align(case["guess"])
        ctx.synchronize()
        out["unsharded_align_warm_ms"] = 1e3 * (time.perf_counter() - t0)
        eu, Hu, bu = gu.linearize(Tg)
        out.update(H_rel_vs_unsharded=float(np.abs(H - Hu).max() / np.abs(Hu).max()), b_rel_vs_unsharded=float(np.abs(b - bu).max() / np.abs(bu).max()),
                   pose_dt_vs_unsharded=float(np.abs(T[:3, 3] - Tu[:3, 3]).max()), iterations_equal=main_iters == gu.last_result["iterations"])
        gu = None
    gs.close()
    gs = None
    ctx.close()
    return out


def leg_c3(torch, rgc, local):
    """config C3: A-LOAM feature extraction + ground fit, batch 1024, 16-beam and 32-beam scans: device time of all
    kernels, and end to end from pinned host scans to pinned feature lists + labels."""
    from rgc_slam_b200 import synth
    from rgc_slam_b200.features import FeatureExtractor
    scene = synth.Scene.make(synth.BASE_SEED + 3000)
    traj = synth.trajectory(20, seed=3)
    out = {}
    ctx = rgc.Context(local)
    for beams, az in ((16, 1800), (32, 900)):
        scans = [synth.lidar_scan(scene, traj[f], n_beams=beams, n_azimuth=az, seed=synth.BASE_SEED + 3000 + f) for f in range(16)]
        B = 1024
        fx = FeatureExtractor(ctx, n_rings=beams, max_scans=B, max_points=int(sum(len(scans[i % 16]) for i in range(B))))
        fx.load([scans[i % 16] for i in range(B)])   # concatenated into the extractor's pinned input buffer (outside the timed region)
        fx.run()
        dev, wall = [], []
        for _ in range(3):
            t0 = time.perf_counter()
            ms = fx.run()
            wall.append(time.perf_counter() - t0)
            dev.append(ms)
        out[f"{beams}-beam"] = {"points_per_scan": float(np.mean([len(s) for s in scans])), "batch": B, "device_ms": float(np.median(dev)),
                                "scans_per_s_device": B / (float(np.median(dev)) * 1e-3), "scans_per_s_e2e": B / float(np.median(wall)),
                                "e2e": "pinned host scans in, pinned feature lists + labels + ground parameters out (H2D and D2H inside)",
                                "h2d_bytes": fx.h2d_bytes, "d2h_bytes": fx.d2h_bytes}
        fx.close()
    ctx.close()
    return out


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist
    import rgc_slam_b200 as rgc

    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pairs = build_workload(rank, args.submap_points)
    ctx = rgc.Context(local)
    ext = torch.cuda.ExternalStream(ctx.stream, device=local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")  # > 126 MB L2

    dev = [dict(src=torch.from_numpy(p["src"]).cuda(local), tgt=torch.from_numpy(p["tgt"]).cuda(local)) for p in pairs]
    pin = [dict(src=torch.from_numpy(p["src"]).pin_memory(), tgt=torch.from_numpy(p["tgt"]).pin_memory()) for p in pairs]
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: device-resident inputs, CUDA events on the library's stream
    leg_cold(torch, rgc, ctx, ext, flush, pairs, dev, max(args.warmup, len(pairs)))
    sampler = ClockSampler(local)
    sampler.prepare()
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    cold = leg_cold(torch, rgc, ctx, ext, flush, pairs, dev, args.steps)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    dev_ms = float(cold["times"].sum())
    t = torch.tensor([dev_ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    value = world * args.steps / (dev_ms_max / 1e3)
    g, T = cold["last"]
    res = g.last_result

    # ---------------- e2e: pinned host clouds through the public call, wall clock
    leg_cold(torch, rgc, ctx, ext, flush, pairs, pin, max(3, args.warmup))
    barrier()
    e2e = leg_cold(torch, rgc, ctx, ext, flush, pairs, pin, args.steps, wall=True)
    t = torch.tensor([float(e2e["times"].sum())], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps / float(t.item())
    n_src, n_tgt = len(pairs[0]["src"]), len(pairs[0]["tgt"])
    h2d = 16 * (n_src + n_tgt)
    d2h = 64 + (res["n_linearize"] * 29 + res["n_compute_error"]) * 8 + 296 * 24 * 2 + 22 * 4 * 2

    log(f"headline legs done: {value:.1f} aligns/s device, {e2e_value:.1f} e2e")
    # ---------------- per-kernel numbers of one C2 sweep (CUDA events recorded by the library around each launch)
    peak, peak_src = peak_hbm()
    gw = new_reg(rgc, ctx)
    gw.setInputTarget(dev[0]["tgt"])
    gw.setInputSource(dev[0]["src"])
    Tg = pairs[0]["guess"].astype(np.float64)
    gw.linearize(Tg)
    ctx.set_profiling(True)
    km = []
    for _ in range(7):
        gw.linearize(Tg)
        gw.compute_error(Tg)
        km.append(ctx.last_kernel_ms())
    ctx.set_profiling(False)
    kc = {kk: float(np.median([x[kk] for x in km])) for kk in km[0]}
    gw = None
    st = cold["stage"]
    it_mean = float(np.mean(cold["iters"]))
    # share of a cold step: one un-hinted (~2x) + it_mean hinted correspondence searches vs everything else
    corr_step_ms = kc["k_correspond"] * (it_mean + 2.0)
    corr_bytes = n_src * (16 + 4 + 4 + 16)  # p in, corr + d2 out, the neighbour found
    kernels = {
        "k_correspond (1-NN, hinted, 1 sweep; 4 lanes per query)": {"ms": kc["k_correspond"], "queries": n_src, "Mqueries_per_s": n_src / kc["k_correspond"] / 1e3, "bound": "sm"},
        "k_knn_tile + k_knn_warp k=20 (source sweep)": {"ms": st["src_knn"], "queries": n_src, "Mqueries_per_s": n_src / max(st["src_knn"], 1e-9) / 1e3, "bound": "sm"},
        "k_linearize (1 sweep)": {"ms": kc["k_linearize"], "bound": "latency at one sweep (4 MB); hbm at batch scale: see hbm_kernels"},
        "k_compute_error (1 sweep)": {"ms": kc["k_compute_error"], "bound": "latency at one sweep"},
        "target build (sort + voxel hash, 500k)": {"ms": st["tgt_build"], "bound": "latency: 9 dependent kernels (keys + histograms, 5 one-launch radix passes, cell counts, tables) and one host wait"},
    }
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    tinfo = json.load(open(tpath)).get("k_correspond", {}) if os.path.exists(tpath) else {}
    roofline = {"bound": "sm", "kernel": "k_correspond", "achieved": corr_bytes / (kc["k_correspond"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": corr_bytes / (kc["k_correspond"] * 1e-3) / 1e9 / peak, "traffic": tinfo.get("bytes"), "peak_source": peak_src,
                "share_of_step": corr_step_ms / (dev_ms_max / args.steps),
                "sm": {kk: tinfo[kk] for kk in tinfo if kk != "bytes"},
                "note": "with on-demand target covariances the largest share of a cold step is the exact 1-NN correspondence search (one launch per "
                        "linearize), a divergent tree walk over an L2-resident voxel hash: SM-issue bound, reported as queries/s and ncu SM throughput "
                        "(profiles/); `achieved` is its algorithmic bytes over its time for completeness.  The HBM-bound kernels of the path "
                        "(covariance, linearize, compute_error) are measured at 8 M / 2 M points under hbm_kernels.",
                "kernels": kernels}
    # ---------------- CPU baseline (oracle port), rank 0, N=1 only, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(pairs, n_aligns=3)
        log("cpu_baseline done")

    # ---------------- the line, as far as the required legs go.  The optional legs below add their keys to it; a
    # watchdog prints it without them if they overrun their budget (a leg stuck inside a library call must not cost
    # the round its headline number)
    Tt = pairs[(args.steps - 1) % len(pairs)]["truth"]
    tt = cold["times"]
    out = {
        "metric": METRIC, "value": value, "unit": "aligns/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "p50_ms": float(np.percentile(tt, 50)),
        "p99_ms": float(np.percentile(tt, 99)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": config_dict(n_src, n_tgt, len(pairs), world),
        "e2e": {"value": e2e_value, "unit": "aligns/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "p50_ms": 1e3 * float(np.percentile(e2e["times"], 50)),
                "p99_ms": 1e3 * float(np.percentile(e2e["times"], 99)),
                "timing": "wall clock around new object + setInputTarget + setInputSource + align with pinned host clouds"},
        "gpu_launches": int(cold["launches"]),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "stage_ms": st,
        "stage_ms_note": "per-stage CUDA-event times on each stage's own stream; the source stages run on a second "
                         "stream concurrently with the target stages, so the stages do not add up to ms_per_step",
        "lm_iterations_mean": it_mean,
        "wall_s_timed_region": wall,
        "pose_err_vs_truth_m": float(np.abs(T[:3, 3] - Tt[:3, 3]).max()),
    }
    extra = {}
    emitted = threading.Lock()

    def emit(note=None):
        if not emitted.acquire(blocking=False):
            return
        if rank == 0:
            line = dict(out)
            line.update(extra)
            if note:
                line["optional_legs"] = note
            print(json.dumps(line), flush=True)

    def overrun():
        log(f"optional legs overran {args.leg_budget:.0f} s: printing the line without the unfinished ones")
        emit(f"stopped after {args.leg_budget:.0f} s; finished: {sorted(extra)}")
        os._exit(0)

    dog = threading.Timer(args.leg_budget, overrun)
    dog.daemon = True
    dog.start()

    if world == 1:
        nshort = max(10, min(args.steps, 40))
        # ---------------- the reference's covariance schedule: all 500k target covariances every frame
        leg_cold(torch, rgc, ctx, ext, flush, pairs, dev, 3, on_demand=False)
        eager = leg_cold(torch, rgc, ctx, ext, flush, pairs, dev, nshort, on_demand=False)
        extra["eager_target_covariances"] = {
            "ms_per_step": float(eager["times"].mean()), "aligns_per_s": 1e3 / float(eager["times"].mean()), "stage_ms": eager["stage"],
            "note": "rgc_reg_set_target_covariance_mode(0): k-NN + covariance of ALL target points at the first align, as fast_gicp_impl.hpp:107-109; "
                    "the default computes them for the points that become correspondences (identical values, tests/test_gpu_round2.py)"}
        # ---------------- warm: target cached (same object), source changes
        gw = new_reg(rgc, ctx)
        gw.setInputTarget(dev[0]["tgt"])
        gw.setInputSource(dev[0]["src"])
        gw.align(pairs[0]["guess"])
        warm = []
        for i in range(nshort):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            gw.setInputSource(dev[0]["src"][:])
            gw.align(pairs[0]["guess"])
            e1.record(ext)
            e1.synchronize()
            warm.append(e0.elapsed_time(e1))
        extra["warm_ms_per_align"] = float(np.mean(warm))
        gw = None
        # ---------------- FastVGICP (what RGC_odometer.cpp:998 instantiates), same workload, cold
        vms, v = [], None
        for i in range(3 + nshort):
            with torch.cuda.stream(ext):
                flush.zero_()
            ctx.synchronize()
            p, cl = pairs[i % len(pairs)], dev[i % len(pairs)]
            v = None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            v = rgc.FastVGICP(ctx)
            v.setResolution(1.0)
            v.setMaximumIterations(CALL_SITE["max_iterations"])
            v.setMaxCorrespondenceDistance(CALL_SITE["corr_dist"])
            v.setTransformationEpsilon(CALL_SITE["transformation_epsilon"])
            v.setInputTarget(cl["tgt"][:])
            v.setInputSource(cl["src"][:])
            v.align(p["guess"])
            e1.record(ext)
            e1.synchronize()
            if i >= 3:
                vms.append(e0.elapsed_time(e1))
        extra["vgicp"] = {"cold_ms_per_align": float(np.mean(vms)), "aligns_per_s": 1e3 / float(np.mean(vms)), "iterations": v.last_result["iterations"],
                          "stage_ms": v.stage_ms(), "params": "resolution 1.0, DIRECT1, ADDITIVE, 25 it, trans_eps 1e-6 (RGC_odometer.cpp:1000-1006)"}
        v = None
        # ---------------- several registrations in flight: T host threads, one context each, pinned host clouds, cold
        if args.concurrent > 1:
            nthr = args.concurrent
            ctxs = [rgc.Context(local) for _ in range(nthr)]
            per_thread = max(4, nshort // 2)

            def worker(tid, n):
                for i in range(n):
                    p, c = pairs[(i + tid) % len(pairs)], pin[(i + tid) % len(pairs)]
                    gg = new_reg(rgc, ctxs[tid])
                    gg.setInputTarget(c["tgt"][:])
                    gg.setInputSource(c["src"][:])
                    gg.align(p["guess"])
                    gg = None

            for tid in range(nthr):
                worker(tid, 2)
            torch.cuda.synchronize()
            th = [threading.Thread(target=worker, args=(tid, per_thread)) for tid in range(nthr)]
            t0 = time.perf_counter()
            for x in th:
                x.start()
            for x in th:
                x.join()
            for cx in ctxs:
                cx.synchronize()
            dtc = time.perf_counter() - t0
            extra["concurrent"] = {"threads": nthr, "aligns": nthr * per_thread, "aligns_per_s": nthr * per_thread / dtc,
                                   "timing": "wall clock, pinned host clouds, cold target every align, one context per host thread"}
            for cx in ctxs:
                cx.close()

        log("eager / warm / vgicp / concurrent legs done")
    ctx.close()
    del dev, flush
    torch.cuda.empty_cache()
    if world == 1 and not args.no_large:
        try:
            roofline["hbm_kernels"] = leg_roofline_large(torch, rgc, local)
        except Exception as e:  # noqa: BLE001 — the headline must still be printed
            roofline["hbm_kernels"] = {"error": f"{type(e).__name__}: {e}"}

        log("hbm_kernels leg done")

    # ---------------- the other configurations
    if not args.no_extra:
        legs = [("c3", lambda: leg_c3(torch, rgc, local)), ("c4", lambda: run_c4(torch, dist, rgc, local, rank, world, args.c4_pairs or 4096))] if world == 1 else \
               [("c4", lambda: run_c4(torch, dist, rgc, local, rank, world, args.c4_pairs or 4096)), ("c5", lambda: run_c5(torch, dist, rgc, local, rank, world))]
        for name, fn in legs:
            try:
                extra[name] = fn()
            except Exception as e:  # noqa: BLE001
                if world > 1:
                    raise  # a collective leg that fails on one rank would hang the others
                extra[name] = {"error": f"{type(e).__name__}: {e}"}

            log(f"{name} leg done")
    dog.cancel()
    emit()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(pairs, n_aligns=3):
    from oracle import oracle as orc
    threads = host_threads()
    secs = []
    for i in range(n_aligns):
        p = pairs[i % len(pairs)]
        o = orc.FastGICP(max_iterations=CALL_SITE["max_iterations"], corr_dist=CALL_SITE["corr_dist"],
                         transformation_epsilon=CALL_SITE["transformation_epsilon"], num_threads=threads)
        t0 = time.perf_counter()
        o.setInputTarget(p["tgt"])   # kd-tree build (fast_gicp_impl.hpp:88)
        o.setInputSource(p["src"])
        o.align(p["guess"])          # lazy covariances + LM (fast_gicp_impl.hpp:103-112)
        secs.append(time.perf_counter() - t0)
    vsecs = []
    for i in range(1):
        p = pairs[i % len(pairs)]
        ov = orc.FastVGICP(resolution=1.0, max_iterations=CALL_SITE["max_iterations"], transformation_epsilon=CALL_SITE["transformation_epsilon"],
                           num_threads=threads)
        t0 = time.perf_counter()
        ov.setInputTarget(p["tgt"])
        ov.setInputSource(p["src"])
        ov.align(p["guess"])
        vsecs.append(time.perf_counter() - t0)
    return {"value": 1.0 / float(np.mean(secs)), "unit": "aligns/s", "cores": threads, "kind": "port", "vgicp_aligns_per_s": 1.0 / float(np.mean(vsecs)),
            "sample": f"{n_aligns} cold aligns of the same C2 pairs (sweep vs {len(pairs[0]['tgt'])}-pt submap), OpenMP guided,8 on all "
                      f"{threads} host threads; the reference call site asks for 14 threads (RGC_odometer.cpp:1006)",
            "seconds_per_align": [float(s) for s in secs]}


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path: the OpenMP FastGICP restatement in oracle/
    (the reference cannot be compiled in this image: no PCL/Eigen/FLANN)."""
    if rank != 0:
        return
    from oracle import oracle as orc
    pairs = build_workload(0, args.submap_points)
    # all host threads this process may use, asked for explicitly: torchrun exports OMP_NUM_THREADS=1
    threads = host_threads()
    # each step is one cold align (~0.2 s on 16 threads): a bounded sample keeps the run within minutes whatever K is
    steps = min(args.steps, 60)
    warmup = min(args.warmup, 3)

    def step(i):
        p = pairs[i % len(pairs)]
        o = orc.FastGICP(max_iterations=CALL_SITE["max_iterations"], corr_dist=CALL_SITE["corr_dist"],
                         transformation_epsilon=CALL_SITE["transformation_epsilon"], num_threads=threads)
        o.setInputTarget(p["tgt"])
        o.setInputSource(p["src"])
        o.align(p["guess"])

    for i in range(warmup):
        step(i)
    times = []
    for i in range(steps):
        t0 = time.perf_counter()
        step(i)
        times.append(time.perf_counter() - t0)
    dt = float(np.sum(times))
    v = steps / dt
    n_src, n_tgt = len(pairs[0]["src"]), len(pairs[0]["tgt"])
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "aligns/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "p50_ms": 1e3 * float(np.percentile(times, 50)),
        "p99_ms": 1e3 * float(np.percentile(times, 99)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": config_dict(n_src, n_tgt, len(pairs), world),
        "cpu_baseline": {"value": v, "unit": "aligns/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} cold aligns (bounded sample of the {args.steps} requested steps), CPU restatement of the reference OpenMP "
                                   f"FastGICP path, {threads} threads"},
        "e2e": {"value": v, "unit": "aligns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--submap-points", type=int, default=N_SUBMAP)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-large", action="store_true", help="skip the 8 M / 2 M HBM-kernel leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the C3 / C4 / C5 legs")
    ap.add_argument("--c4-pairs", type=int, default=0, help="pairs of the C4 leg (default 4096 at every N: one strong-scaling series)")
    ap.add_argument("--leg-budget", type=float, default=420.0, help="seconds the optional legs (everything after the headline, per-kernel and "
                    "cpu_baseline numbers) may take before the line is printed without the unfinished ones")
    ap.add_argument("--concurrent", type=int, default=4, help="host threads of the concurrent-throughput leg (0/1 = skip)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
