# round 2, GPU call 8: parity + all-thread final pass / fitness hints / chunk pipeline effects
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2c8_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c8_pytest.log
tail -4 gpurun_out/r2c8_pytest.log
timeout 300 python tools/roofline_large.py 16 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('large', {k: (round(d[k]['ms'],4), round(d[k].get('frac_of_peak',0),3)) for k in ('k_covariance','k_linearize','k_compute_error','k_correspond','k_knn_tile')})"
timeout 600 python tools/bench_c4.py --pairs 512 --batched > gpurun_out/r2c8_c4.json 2>gpurun_out/r2c8_c4.err; tail -c 700 gpurun_out/r2c8_c4.json
timeout 600 python bench.py --steps 100 --warmup 16 --no-cpu --no-large --no-extra --concurrent 0 2>/dev/null | tail -1 > gpurun_out/r2c8_bench.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2c8_bench.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "p50", d["p50_ms"], "p99", d["p99_ms"], "e2e", d["e2e"]["value"], "iters", d["lm_iterations_mean"])
print("stage", d["stage_ms"]); print({k: v["ms"] for k, v in d["roofline"]["kernels"].items()})
PY
