# round 2, GPU call 22: source self-kNN variants (all-warp, lower deferral threshold), timeline at the new defaults
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gicp.py -m gpu -x -q > gpurun_out/r2c22_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c22_pytest.log
RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_vB.so timeout 600 python -m pytest tests/test_gpu_gicp.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/r2c22_pytest_allwarp.log 2>&1; echo "pytest allwarp rc=$?"; tail -3 gpurun_out/r2c22_pytest_allwarp.log
run() { timeout 300 python bench.py --steps 60 --warmup 8 --no-cpu --no-large --no-extra --concurrent 0 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), 'p50', round(d['p50_ms'],3), 'e2e ms', round(1e3/d['e2e']['value'],3), 'warm', round(d['warm_ms_per_align'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
for r in 1 2; do
  run new
  RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_vB.so run allwarp
  RGC_KNN_DEFER=60 run defer60
  RGC_KNN_DEFER=300 run defer300
done 2>&1 | tee gpurun_out/r2c22_ab.txt
RGC_TIMELINE=1 timeout 300 python tools/prof_step.py 6 > gpurun_out/r2c22_timeline.log 2> gpurun_out/r2c22_timeline.err
python - <<'PY'
lines = open("gpurun_out/r2c22_timeline.err").read().split("\n")
starts = [i for i, l in enumerate(lines) if "marks (us since" in l]
print("\n".join(lines[starts[-1]:]))
PY
