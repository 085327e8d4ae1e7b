# round 2, GPU call 6: parity suite + the full default bench line (timing of the whole run)
mkdir -p gpurun_out
SECONDS=0
timeout 900 python bench.py > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err; echo "bench rc=$? wall ${SECONDS}s"; tail -3 gpurun_out/r2c6_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2c6_bench.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "p50", d["p50_ms"], "p99", d["p99_ms"], "e2e", d["e2e"]["value"], "iters", d["lm_iterations_mean"])
print("stage", d["stage_ms"])
print("eager", d["eager_target_covariances"]["ms_per_step"], "warm", d["warm_ms_per_align"], "vgicp", d["vgicp"]["cold_ms_per_align"], "conc", d.get("concurrent",{}).get("aligns_per_s"))
print("roofline", {k: d["roofline"][k] for k in ("kernel","achieved","frac","share_of_step")})
hk=d["roofline"].get("hbm_kernels",{})
for k in ("k_covariance","k_linearize","k_compute_error"):
    if k in hk: print(k, {kk: round(v,3) if isinstance(v,float) else v for kk,v in hk[k].items()})
print("c3", d.get("c3")); print("c4", d.get("c4")); print("cpu", d.get("cpu_baseline"))
PY
