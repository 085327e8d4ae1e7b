# round 2, GPU call 24: correspondence kernels at 10 blocks/SM (48 registers) so that 8 lanes x 22.7k queries fit one wave
mkdir -p gpurun_out
RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_vB.so timeout 600 python -m pytest tests/test_gpu_gicp.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/r2c24_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c24_pytest.log
run() { timeout 300 python bench.py --steps 60 --warmup 8 --no-cpu --no-large --no-extra --concurrent 0 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), 'p50', round(d['p50_ms'],3), 'e2e ms', round(1e3/d['e2e']['value'],3), 'warm', round(d['warm_ms_per_align'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
for r in 1 2 3; do
  run minb8
  RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_vB.so run minb10
done 2>&1 | tee gpurun_out/r2c24_ab.txt
