# round 2, GPU call 21: full parity run at the new defaults (8 lanes per query, batch merge), bench, C4, N=1 default line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c21_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c21_pytest.log
tail -6 gpurun_out/r2c21_pytest.log
timeout 1200 python bench.py > gpurun_out/r2c21_bench_default.json 2> gpurun_out/r2c21_bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c21_bench_default.json").read().strip().split("\n")[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "p50", d["p50_ms"], "p99", d["p99_ms"], "e2e", d["e2e"]["value"], "iters", d["lm_iterations_mean"])
print("stage", d["stage_ms"]); print("eager", d.get("eager_target_covariances", {}).get("ms_per_step"), "warm", d.get("warm_ms_per_align"), "vgicp", d.get("vgicp", {}).get("cold_ms_per_align"), "conc", d.get("concurrent"))
hk = d["roofline"].get("hbm_kernels", {})
print({k: (round(v["ms"], 4), round(v.get("frac_layout_bytes", 0), 3), round(v.get("frac_survey_bytes", 0), 3)) for k, v in hk.items() if isinstance(v, dict) and "ms" in v})
print("c3", d.get("c3")); print("c4", d.get("c4")); print("cpu", d.get("cpu_baseline"))
PY
