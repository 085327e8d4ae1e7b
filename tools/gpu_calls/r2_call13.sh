# round 2, GPU call 13: four-node DFS steps in k_knn_warp, 8-wide look-back; variants: root level + 1, k_knn_warp at 3 blocks/SM
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gicp.py tests/test_gpu_fullsize.py tests/test_gpu_round2.py tests/test_gpu_batch.py -m gpu -x -q > gpurun_out/r2c13_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c13_pytest.log
tail -4 gpurun_out/r2c13_pytest.log
run() { timeout 300 python bench.py --steps 60 --warmup 8 --no-cpu --no-large --no-extra --concurrent 0 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), 'p50', round(d['p50_ms'],3), 'e2e ms', round(1e3/d['e2e']['value'],3), 'warm', round(d['warm_ms_per_align'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
for r in 1 2; do
  RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_prev.so run prev
  run new
  RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_vB.so run rootshift1
  RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_vC.so run kwminb3
done 2>&1 | tee gpurun_out/r2c13_ab.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2c13_launches.csv python bench.py --steps 2 --warmup 2 --no-cpu --no-large --no-extra --concurrent 0 > gpurun_out/r2c13_bench_under_ncu.log 2>&1
