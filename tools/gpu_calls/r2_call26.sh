# round 2, GPU call 26: main stream joins the source lane behind its SORT (not its tables); radix sort tile size / look-back width
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gicp.py tests/test_gpu_round2.py tests/test_gpu_batch.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2c26_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c26_pytest.log
tail -4 gpurun_out/r2c26_pytest.log
RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_vA.so timeout 600 python -m pytest tests/test_gpu_gicp.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
run() { timeout 300 python bench.py --steps 60 --warmup 8 --no-cpu --no-large --no-extra --concurrent 0 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), 'p50', round(d['p50_ms'],3), 'e2e ms', round(1e3/d['e2e']['value'],3), 'warm', round(d['warm_ms_per_align'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
for r in 1 2; do
  RGC_JOIN_TABLES=1 run join_tables
  run join_sort
  RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_vA.so run items4
  RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_vB.so run items4_lb16
  RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_vC.so run items6
done 2>&1 | tee gpurun_out/r2c26_ab.txt
RGC_TIMELINE=1 timeout 300 python tools/prof_step.py 6 > gpurun_out/r2c26_timeline.log 2> gpurun_out/r2c26_timeline.err
python - <<'PY'
lines = open("gpurun_out/r2c26_timeline.err").read().split("\n")
starts = [i for i, l in enumerate(lines) if "marks (us since" in l]
print("\n".join(lines[starts[-1]:starts[-1]+22]))
PY
