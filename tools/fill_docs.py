"""Regenerate the measured-number blocks of DESIGN.md / README.md / profiles/README.md from the bench
lines committed under profiles/ (so the prose never drifts from the JSON).  Blocks are delimited by
`<!-- NAME:BEGIN -->` / `<!-- NAME:END -->`; a bare `@@NAME@@` token is upgraded to such a block.

    python tools/fill_docs.py [profiles/r2_bench_default_1gpu.json [profiles/r2_bench_reference_arm.json]]
"""
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def load(p):
    return json.loads(Path(p).read_text().strip().split("\n")[-1])


def block(text, name, body):
    begin, end = f"<!-- {name}:BEGIN -->", f"<!-- {name}:END -->"
    new = f"{begin}\n{body.rstrip()}\n{end}"
    if f"@@{name}@@" in text:
        return text.replace(f"@@{name}@@", new)
    pat = re.compile(re.escape(begin) + r".*?" + re.escape(end), re.S)
    if not pat.search(text):
        raise SystemExit(f"marker {name} not found")
    return pat.sub(lambda m: new, text)


def main():
    g = load(sys.argv[1] if len(sys.argv) > 1 else ROOT / "profiles/r2_bench_default_1gpu.json")
    r = load(sys.argv[2] if len(sys.argv) > 2 else ROOT / "profiles/r2_bench_reference_arm.json")
    e2e, ref = g["e2e"], r["value"]
    hk = g["roofline"]["hbm_kernels"]
    eager = g["eager_target_covariances"]
    c3, c4 = g["c3"], g["c4"]

    def f(k, which):
        return 100.0 * hk[k][f"frac_{which}_bytes"]

    headline = (
        f"**{g['ms_per_step']:.3f} ms per cold align (p50 {g['p50_ms']:.3f}, p99 {g['p99_ms']:.2f}) = {g['value']:.0f} aligns/s** "
        f"with the clouds already in HBM; **{e2e['value']:.0f} aligns/s end to end** (p50 {e2e['p50_ms']:.3f} ms) through "
        f"`FastGICP.setInputTarget / setInputSource / align` from pinned host clouds ({e2e['h2d_bytes_per_step'] / 1e6:.2f} MB H2D + "
        f"{e2e['d2h_bytes_per_step'] / 1e3:.1f} kB D2H inside the timed region); warm target (sweep-to-same-map) "
        f"{g['warm_ms_per_align']:.3f} ms.  The CPU restatement of the reference on the box's {r['cpu_baseline']['cores']} host threads: "
        f"{ref:.2f} aligns/s ({1e3 / ref:.0f} ms) ⇒ **{e2e['value'] / ref:.0f}× end to end**, {g['value'] / ref:.0f}× device-timed.  "
        f"Reference schedule (k-NN + covariance of all 500k target points every frame, `rgc_reg_set_target_covariance_mode(0)`): "
        f"{eager['ms_per_step']:.3f} ms; FastVGICP (DIRECT1, resolution 1.0): {g['vgicp']['cold_ms_per_align']:.3f} ms.  "
        f"Clocks {g['clocks']['sm_mhz']:.0f}/{g['clocks']['sm_max_mhz']:.0f} MHz, throttle reasons {g['clocks']['reasons']}.\n\n"
        f"HBM kernels at batch scale (8 M target / 2 M source points, fractions of the measured {hk['peak_GBps']:.0f} GB/s): "
        f"`k_covariance` {hk['k_covariance']['ms']:.3f} ms = {f('k_covariance', 'survey'):.0f} % on SURVEY §8(d) bytes / "
        f"{f('k_covariance', 'layout'):.0f} % on the bytes this layout moves; `k_linearize` {hk['k_linearize']['ms']:.3f} ms = "
        f"{f('k_linearize', 'survey'):.0f} % / {f('k_linearize', 'layout'):.0f} %; `k_compute_error` {hk['k_compute_error']['ms']:.3f} ms = "
        f"{f('k_compute_error', 'survey'):.0f} % / {f('k_compute_error', 'layout'):.0f} %.\n\n"
        f"Feature extraction (C3, batch 1024): 16-beam {c3['16-beam']['scans_per_s_device']:.0f} scans/s on the device, "
        f"{c3['16-beam']['scans_per_s_e2e']:.0f} end to end (pinned in, pinned lists + labels out); 32-beam "
        f"{c3['32-beam']['scans_per_s_device']:.0f} / {c3['32-beam']['scans_per_s_e2e']:.0f}.  Batched loop-closure verification "
        f"(C4, `rgc_batch_align`, host clouds in): {c4['pairs_per_s']:.0f} pairs/s on one GPU, {c4['recovered_truth']}/{c4['pairs']} recover the truth."
    )
    glance = (
        "| what | number |\n|---|---|\n"
        f"| C2 cold align, device-timed (200 steps, 16 pairs, L2 flushed) | **{g['ms_per_step']:.3f} ms** mean, p50 {g['p50_ms']:.3f}, p99 {g['p99_ms']:.2f} ({g['value']:.0f} aligns/s) |\n"
        f"| C2 end to end from pinned host clouds | {e2e['value']:.0f} aligns/s (p50 {e2e['p50_ms']:.3f} ms) |\n"
        f"| CPU restatement, {r['cpu_baseline']['cores']} threads (`--impl reference`) | {ref:.2f} aligns/s ⇒ {e2e['value'] / ref:.0f}× e2e, {g['value'] / ref:.0f}× device |\n"
        f"| warm align (target kept) | {g['warm_ms_per_align']:.3f} ms |\n"
        f"| eager schedule (all 500k target covariances) | {eager['ms_per_step']:.3f} ms (`k_knn_tile` {eager['stage_ms']['tgt_knn']:.3f}) |\n"
        f"| FastVGICP cold | {g['vgicp']['cold_ms_per_align']:.3f} ms |\n"
        f"| 4 host threads, one context each | {g['concurrent']['aligns_per_s']:.0f} aligns/s |\n"
        f"| stages (own streams) | src build {g['stage_ms']['src_build']:.3f}, src k-NN {g['stage_ms']['src_knn']:.3f}, src cov {g['stage_ms']['src_cov']:.3f}, tgt build {g['stage_ms']['tgt_build']:.3f}, LM {g['stage_ms']['lm']:.3f} ms ({g['lm_iterations_mean']:.1f} iterations) |\n"
        f"| `k_covariance` 8 M | {hk['k_covariance']['ms']:.3f} ms, {f('k_covariance', 'survey'):.0f} % (§8d bytes) / {f('k_covariance', 'layout'):.0f} % (layout bytes) of HBM |\n"
        f"| `k_linearize` 2 M | {hk['k_linearize']['ms']:.3f} ms, {f('k_linearize', 'survey'):.0f} % / {f('k_linearize', 'layout'):.0f} % |\n"
        f"| `k_compute_error` 2 M | {hk['k_compute_error']['ms']:.3f} ms, {f('k_compute_error', 'survey'):.0f} % / {f('k_compute_error', 'layout'):.0f} % |\n"
        f"| C3 features, batch 1024 | 16-beam {c3['16-beam']['scans_per_s_device']:.0f} scans/s device, {c3['16-beam']['scans_per_s_e2e']:.0f} e2e; 32-beam {c3['32-beam']['scans_per_s_device']:.0f} / {c3['32-beam']['scans_per_s_e2e']:.0f} |\n"
        f"| C4 batched pairs, one GPU | {c4['pairs_per_s']:.0f} pairs/s ({c4['pairs']} pairs, host clouds in) |"
    )
    last_row = f"{g['ms_per_step']:.3f} | {g['p50_ms']:.3f} | {1e3 / e2e['value']:.3f} | {g['warm_ms_per_align']:.2f}"

    for path, name, body in (("DESIGN.md", "DESIGN_HEADLINE", headline), ("README.md", "README_HEADLINE", "Measured on one B200 (`profiles/r2_bench_default_1gpu.json`): " + headline),
                             ("profiles/README.md", "R2_GLANCE", glance)):
        p = ROOT / path
        p.write_text(block(p.read_text(), name, body))
    p = ROOT / "profiles/README.md"
    t = p.read_text()
    if "@@R2_LAST_ROW@@" in t:
        t = t.replace("@@R2_LAST_ROW@@", last_row)
    p.write_text(t)
    print("filled DESIGN.md, README.md, profiles/README.md")


if __name__ == "__main__":
    main()
