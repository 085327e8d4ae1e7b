# round 2, GPU call 11: late join of the source covariances + compute_error fused into the look-ahead correspondence launch
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c11_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c11_pytest.log
tail -4 gpurun_out/r2c11_pytest.log
timeout 300 python tools/need_stats.py 3 2>&1 | tail -4 | tee gpurun_out/r2c11_need_stats.txt
run() { timeout 300 python bench.py --steps 60 --warmup 8 --no-cpu --no-large --no-extra --concurrent 0 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), 'p50', round(d['p50_ms'],3), 'e2e ms', round(1e3/d['e2e']['value'],3), 'warm', round(d['warm_ms_per_align'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
for r in 1 2; do
  RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_prev.so run prev
  run new
  RGC_NO_FUSE_TRIAL=1 run new_nofuse
  RGC_NO_LATE_JOIN=1 run new_nolatejoin
done 2>&1 | tee gpurun_out/r2c11_ab.txt
