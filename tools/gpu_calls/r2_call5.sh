# round 2, GPU call 5: parity suite (canonical grid, deferred builds, batch) + headline + C4
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2c5_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c5_pytest.log
tail -6 gpurun_out/r2c5_pytest.log
run() { tag=$1; shift; env "$@" timeout 400 python bench.py --steps 30 --warmup 4 --no-cpu --concurrent 0 2> gpurun_out/r2c5_$tag.err | tail -1 > gpurun_out/r2c5_$tag.json; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2c5_$tag.json"))
    print("$tag", "ms/step", round(d["ms_per_step"],3), "e2e ms", round(1e3/d["e2e"]["value"],3), "warm", round(d["warm_ms_per_align"],3), "vgicp", round(d["vgicp"]["cold_ms_per_align"],3), {k: round(v,3) for k,v in d["stage_ms"].items()}, "launches", d["gpu_launches"])
except Exception as e:
    print("$tag failed", e)
PY
}
run default A=1
run syncbuild RGC_SYNC_BUILD=1
run eager RGC_EAGER_TARGET_COV=1
timeout 600 python tools/bench_c4.py --pairs 256 --batched > gpurun_out/r2c5_c4_batched.json 2> gpurun_out/r2c5_c4_batched.err; tail -c 900 gpurun_out/r2c5_c4_batched.json; tail -3 gpurun_out/r2c5_c4_batched.err
