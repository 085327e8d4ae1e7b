"""Turn an .ncu-rep (from gpurun_out/) into a small, committable text summary under profiles/.
    python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep profiles/X_summary.txt [kernel-regex-for-source-hotspots]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    hot = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none summary of {rep}", "# (times are cold-cache and serialised under the profiler: compare shares, not absolutes)", ""]
    stall_cols = [h for h in hdr if "issue_stalled" in h and "per_warp_active" in h]
    for r in rows[2:]:
        lines.append("kernel: " + r[hdr.index("Kernel Name")][:110])
        for k in KEYS:
            if k in hdr:
                lines.append(f"  {k:72s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
        st = sorted(((float(r[hdr.index(s)].replace(",", "") or 0), s) for s in stall_cols), reverse=True)[:4]
        for v, s in st:
            lines.append(f"  top stall  {s.replace('smsp__average_warp_latency_', '').replace('smsp__average_warps_', '')[:60]:60s} {v:10.2f}")
        lines.append("")
    if hot:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{hot}",
                              "--launch-skip", sys.argv[4] if len(sys.argv) > 4 else "0", "--launch-count", "1"], capture_output=True, text=True).stdout
        cur, agg, tot = None, [], 0
        for r in csv.reader(io.StringIO(src)):
            if len(r) >= 2 and r[0] == "File Path":
                cur = r[1].split("/")[-1]
            elif len(r) > 8 and r[0].isdigit():
                try:
                    v, t, s = int(r[7]), int(r[8]), int(r[4])
                except ValueError:
                    continue
                agg.append((v, t, s, cur, int(r[0]), r[1].strip()[:90]))
                tot += v
        agg.sort(reverse=True)
        ts = sum(a[2] for a in agg) or 1
        lines.append(f"source hot spots of {hot} (warp instructions executed: {tot}):")
        for v, t, s, f, l, text in agg[:25]:
            lines.append(f"  {100 * v / max(tot, 1):5.1f}% instr {100 * s / ts:5.1f}% stall-samples  active lanes {t / max(v, 1):5.1f}  {f}:{l}  {text}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
