#!/usr/bin/env python
"""bench.py — scan-to-map GICP aligns/s (VLP-16 sweep vs 500k-point submap), BASELINE.json config C2.

A "step" is what the odometry node does per LiDAR frame (rgc_slam/src/RGC_odometer.cpp:998-1011):
a fresh FastGICP object, the call-site parameters (25 iterations, corr 2 m, trans eps 1e-6),
setInputTarget(NEW 500k-point submap) + setInputSource(new sweep) + align(guess).  The target
changes every frame at the reference call site, so every step is COLD: target voxel-hash build,
k=20 kNN and covariances are inside the timed region (`warm` = target cached is reported beside it).

  value : aligns/s, inputs already resident in HBM, device time (CUDA events on the library's
          stream, summed over the K steps; L2 flushed between steps outside the events)
  e2e   : aligns/s through the same public call with pinned HOST clouds: H2D of both clouds and
          the D2H of the result are inside the timed (wall-clock) region
  --impl reference : the CPU restatement of the reference's OpenMP FastGICP path (oracle/, the
          reference itself cannot be built here) on the same workload, all host threads

N > 1 (torchrun): independent registrations sharded across ranks, no collective in the data path
(weak scaling); value = total aligns / max-over-ranks time.
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

from rgc_slam_b200 import synth  # noqa: E402

N_SUBMAP = 500_000
CALL_SITE = dict(max_iterations=25, corr_dist=2.0, transformation_epsilon=1e-6)  # RGC_odometer.cpp:1000-1006
N_PAIRS = 2  # distinct (sweep, submap) pairs cycled through


def build_workload(rank: int, n_submap: int, n_pairs: int = N_PAIRS):
    """config C2 stream (rgc_slam_b200/workloads.py): sweeps along a trajectory vs the accumulated submap"""
    from rgc_slam_b200 import workloads
    return workloads.build_c2_pairs(rank, n_submap, n_pairs)


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region.  In-process NVML at 5 Hz (an
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


def new_reg(rgc, ctx):
    g = rgc.FastGICP(ctx)
    g.setMaximumIterations(CALL_SITE["max_iterations"])
    g.setMaxCorrespondenceDistance(CALL_SITE["corr_dist"])
    g.setTransformationEpsilon(CALL_SITE["transformation_epsilon"])
    g.setEuclideanFitnessEpsilon(1e-6)
    g.setRANSACIterations(0)
    g.setNumThreads(14)
    return g


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

    def step(i, clouds, fresh=True):
        p = pairs[i % len(pairs)]
        c = clouds[i % len(pairs)]
        g = new_reg(rgc, ctx)
        # a new tensor view per step = a new cloud identity, as at the reference call site
        g.setInputTarget(c["tgt"][:] if fresh else c["tgt"])
        g.setInputSource(c["src"][:] if fresh else c["src"])
        T = g.align(p["guess"])
        return g, T

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: device-resident inputs, CUDA events on the library's stream
    g = None
    for i in range(max(args.warmup, 2 * len(pairs))):
        g = None  # the reference's object is stack-local: destroyed before the next frame's is built
        g, _ = step(i, dev)
    stage_acc = {}
    iters = []
    sampler = ClockSampler(local)
    sampler.prepare()
    barrier()
    sampler.start()
    dev_ms = 0.0
    wall0 = time.perf_counter()
    n_launch_timed = 0
    for i in range(args.steps):
        with torch.cuda.stream(ext):
            flush.zero_()  # L2 flush between timed iterations (outside the event pair)
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count
        g = None
        e0.record(ext)
        g, T = step(i, dev)
        e1.record(ext)
        e1.synchronize()
        n_launch_timed += ctx.launch_count - l0
        dev_ms += e0.elapsed_time(e1)
        for k, v in g.stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
        iters.append(g.last_result["iterations"])
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    t = torch.tensor([dev_ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    value = world * args.steps / (dev_ms_max / 1e3)

    # ---------------- warm: target cached (same object), source changes
    gw = new_reg(rgc, ctx)
    gw.setInputTarget(dev[0]["tgt"])
    gw.setInputSource(dev[0]["src"])
    gw.align(pairs[0]["guess"])
    warm_ms = 0.0
    nwarm = max(3, min(args.steps, 10))
    for i in range(nwarm):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        gw.setInputSource(dev[0]["src"][:])
        gw.align(pairs[0]["guess"])
        e1.record(ext)
        e1.synchronize()
        warm_ms += e0.elapsed_time(e1)
    warm_ms /= nwarm

    # ---------------- e2e: pinned host clouds through the public call, wall clock
    g = None
    for i in range(max(2, args.warmup // 2)):
        g = None
        g, _ = step(i, pin)
    barrier()
    e2e_s = 0.0
    for i in range(args.steps):
        with torch.cuda.stream(ext):
            flush.zero_()
        ctx.synchronize()
        g = None
        t0 = time.perf_counter()
        g, T = step(i, pin)
        ctx.synchronize()
        e2e_s += time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps / float(t.item())
    n_src, n_tgt = len(pairs[0]["src"]), len(pairs[0]["tgt"])
    h2d = 16 * (n_src + n_tgt)
    res = g.last_result
    d2h = 64 + (res["n_linearize"] * 29 + res["n_compute_error"]) * 8 + 296 * 24 * 2 + 22 * 4 * 2

    # ---------------- FastVGICP (what RGC_odometer.cpp:998 instantiates), same workload, cold
    vg = None
    if world == 1:
        def vstep(i):
            p, cl = pairs[i % len(pairs)], dev[i % len(pairs)]
            v = rgc.FastVGICP(ctx)
            v.setResolution(1.0)
            v.setMaximumIterations(CALL_SITE["max_iterations"])
            v.setMaxCorrespondenceDistance(CALL_SITE["corr_dist"])
            v.setTransformationEpsilon(CALL_SITE["transformation_epsilon"])
            v.setInputTarget(cl["tgt"][:])
            v.setInputSource(cl["src"][:])
            return v, v.align(p["guess"])
        v = None
        for i in range(3):
            v = None
            v, _ = vstep(i)
        vms = 0.0
        nv = max(3, min(args.steps, 10))
        for i in range(nv):
            with torch.cuda.stream(ext):
                flush.zero_()
            ctx.synchronize()
            v = None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            v, Tv = vstep(i)
            e1.record(ext)
            e1.synchronize()
            vms += e0.elapsed_time(e1)
        vg = {"cold_ms_per_align": vms / nv, "aligns_per_s": nv / (vms * 1e-3), "iterations": v.last_result["iterations"],
              "stage_ms": v.stage_ms(), "params": "resolution 1.0, DIRECT1, ADDITIVE, 25 it, trans_eps 1e-6 (RGC_odometer.cpp:1000-1006)"}
        v = None

    # ---------------- throughput with several registrations in flight (SURVEY §8e C4 pattern on the C2
    # workload): T host threads, each with its own context (stream pair + pool), pinned host clouds,
    # cold target every align.  The sequential `value` / `e2e` above are latency-bound (one LM loop at
    # a time leaves most SMs idle); this is what a batch consumer (loop-closure verification, bag
    # replay) gets from one GPU.
    conc = None
    if world == 1 and args.concurrent > 1:
        import threading
        nthr = args.concurrent
        ctxs = [rgc.Context(local) for _ in range(nthr)]
        per_thread = max(4, args.steps // 2)

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
        conc = {"threads": nthr, "aligns": nthr * per_thread, "aligns_per_s": nthr * per_thread / dtc,
                "timing": "wall clock, pinned host clouds, cold target every align, one context per host thread"}
        for cx in ctxs:
            cx.close()

    # ---------------- per-kernel roofline (live CUDA-event stage times from the library)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    k = 20
    st = {kk: v / args.steps for kk, v in stage_acc.items()}
    knn_bytes = n_tgt * (16 + 4 * k)          # read own float4, write k int32 positions
    cov_bytes = n_tgt * (16 + 4 * k + 48)     # + 6 fp64 out (fp64 covariances)
    # one linearize / compute_error call timed alone (launch + sync), and the kernels inside it
    # (CUDA events recorded by the library around each launch: rgc_ctx_set_profiling)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Tg = pairs[0]["guess"].astype(np.float64)
    gw.linearize(Tg)
    lin_ms = ce_ms = 0.0
    for _ in range(5):
        e0.record(ext); gw.linearize(Tg); e1.record(ext); e1.synchronize(); lin_ms += e0.elapsed_time(e1) / 5
        e0.record(ext); gw.compute_error(Tg); e1.record(ext); e1.synchronize(); ce_ms += e0.elapsed_time(e1) / 5
    ctx.set_profiling(True)
    km = []
    for _ in range(5):
        gw.linearize(Tg)
        gw.compute_error(Tg)
        km.append(ctx.last_kernel_ms())
    ctx.set_profiling(False)
    kc = {kk: float(np.median([x[kk] for x in km])) for kk in km[0]}
    lin_bytes = n_src * (16 + 48 + 4 + 16 + 48 + 48)   # p, C_A, corr, q + C_B gathers, M out
    ce_bytes = n_src * (16 + 4 + 16 + 48)

    def gbs(b, ms):
        return b / (ms * 1e-3) / 1e9 if ms > 0 else 0.0

    kernels = {
        "k_knn_tile k=20 (target)": {"ms": st["tgt_knn"], "alg_bytes": knn_bytes, "GBps": gbs(knn_bytes, st["tgt_knn"]), "bound": "sm (see profiles/)"},
        "k_covariance (target)": {"ms": st["tgt_cov"], "alg_bytes": cov_bytes, "GBps": gbs(cov_bytes, st["tgt_cov"]), "bound": "hbm"},
        "k_correspond (1-NN, 1 sweep)": {"ms": kc["k_correspond"], "queries": n_src, "bound": "sm/latency"},
        "k_linearize (1 sweep)": {"ms": kc["k_linearize"], "alg_bytes": lin_bytes, "GBps": gbs(lin_bytes, kc["k_linearize"]),
                                  "bound": "latency at 1 sweep (4 MB); hbm at batch scale: see profiles/README.md"},
        "k_compute_error (1 sweep)": {"ms": kc["k_compute_error"], "alg_bytes": ce_bytes, "GBps": gbs(ce_bytes, kc["k_compute_error"]),
                                      "bound": "latency at 1 sweep"},
        "linearize() call incl. launch+sync": {"ms": lin_ms}, "compute_error() call incl. launch+sync": {"ms": ce_ms},
    }
    dom = "k_knn_tile k=20 (target)"
    traffic, sm_side = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        ent = json.load(open(tpath)).get(dom, {})
        traffic = ent.get("bytes")
        sm_side = {k: ent[k] for k in ("sm_issue_active_pct", "sm_throughput_pct", "warp_instructions", "active_lanes_per_instruction", "sm_source") if k in ent}
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["GBps"] / peak, "traffic": traffic, "peak_source": peak_src, "sm": sm_side,
                "note": "dominant kernel is the k=20 kNN, which is SM/latency-bound (north_star: report SM throughput); "
                        "HBM-bound kernels are listed under `kernels`; ncu summaries in profiles/",
                "kernels": kernels}

    # ---------------- CPU baseline (oracle port), rank 0, N=1 only, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(pairs, n_aligns=2)

    if rank == 0:
        Tt = pairs[(args.steps - 1) % len(pairs)]["truth"]
        out = {
            "metric": "scan-to-map GICP aligns/sec (VLP-16 vs 500k-pt map)", "value": value, "unit": "aligns/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 (H/b, covariances) + f32 (points, kNN distances)", "data": "synthetic",
            "config": {"workload": "C2: VLP-16 sweep vs 500k-point submap, cold target every step (new submap per frame as at "
                                   "RGC_odometer.cpp:985-1009), call-site params 25 it / corr 2 m / trans_eps 1e-6, k=20 PLANE LM",
                       "n_source": n_src, "n_target": n_tgt, "pairs_cycled": len(pairs), "l2": "flushed between steps (256 MB memset)",
                       "sharding": "independent registrations per rank, no collective" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": "aligns/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "timing": "wall clock around setInputTarget+setInputSource+align with pinned host clouds"},
            "gpu_launches": int(n_launch_timed),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "warm_ms_per_align": warm_ms,
            "concurrent": conc,
            "vgicp": vg,
            "stage_ms": st,
            "stage_ms_note": "per-stage CUDA-event times on each stage's own stream; the source stages run on a second "
                             "stream concurrently with the target stages, so the stages do not add up to ms_per_step",
            "lm_iterations_mean": float(np.mean(iters)),
            "wall_s_timed_region": wall,
            "pose_err_vs_truth_m": float(np.abs(T[:3, 3] - Tt[:3, 3]).max()),
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(pairs, n_aligns=2):
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

    def step(i):
        p = pairs[i % len(pairs)]
        o = orc.FastGICP(max_iterations=CALL_SITE["max_iterations"], corr_dist=CALL_SITE["corr_dist"],
                         transformation_epsilon=CALL_SITE["transformation_epsilon"], num_threads=threads)
        o.setInputTarget(p["tgt"])
        o.setInputSource(p["src"])
        o.align(p["guess"])

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    v = args.steps / dt
    n_src, n_tgt = len(pairs[0]["src"]), len(pairs[0]["tgt"])
    out = {
        "impl": "reference", "metric": "scan-to-map GICP aligns/sec (VLP-16 vs 500k-pt map)", "value": v, "unit": "aligns/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 (H/b, covariances) + f32 (points, kNN distances)", "data": "synthetic",
        "config": {"workload": "C2: VLP-16 sweep vs 500k-point submap, cold target every step, call-site params 25 it / corr 2 m / "
                               "trans_eps 1e-6, k=20 PLANE LM", "n_source": n_src, "n_target": n_tgt, "pairs_cycled": len(pairs)},
        "cpu_baseline": {"value": v, "unit": "aligns/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} cold aligns, CPU restatement of the reference OpenMP FastGICP path, {threads} threads"},
        "e2e": {"value": v, "unit": "aligns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--submap-points", type=int, default=N_SUBMAP)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
