# round 2, GPU call 1: parity suite + lazy/eager + A/B of the search variants (one B200)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c1_pytest.log
tail -5 gpurun_out/r2c1_pytest.log
run() { tag=$1; shift; env "$@" timeout 400 python bench.py --steps 30 --warmup 4 --no-cpu --concurrent 0 2> gpurun_out/r2c1_$tag.err | tail -1 > gpurun_out/r2c1_$tag.json; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2c1_$tag.json"))
    print("$tag", "ms/step", round(d["ms_per_step"],3), "e2e ms", round(1e3/d["e2e"]["value"],3), "warm", round(d["warm_ms_per_align"],3), "vgicp", round(d["vgicp"]["cold_ms_per_align"],3), {k: round(v,3) for k,v in d["stage_ms"].items()}, "launches", d["gpu_launches"])
except Exception as e:
    print("$tag failed", e)
PY
}
V=$PWD/rgc_slam_b200/variants
run default A=1
run eager RGC_EAGER_TARGET_COV=1
run nospin RGC_NO_SPIN=1
run nobatch_mb8 RGC_LIB=$V/nobatch_mb8.so
run batch_mb4 RGC_LIB=$V/batch_mb4.so
run batch_mb8 RGC_LIB=$V/batch_mb8.so
run default2 A=1
run nolookahead RGC_NO_LOOKAHEAD=1
