# round 2, GPU call 19: full parity run (k > 32, feature clouds), incremental root keys, k_knn_warp seed count, C4 batched
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c19_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c19_pytest.log
tail -6 gpurun_out/r2c19_pytest.log
run() { timeout 300 python bench.py --steps 60 --warmup 8 --no-cpu --no-large --no-extra --concurrent 0 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), 'p50', round(d['p50_ms'],3), 'e2e ms', round(1e3/d['e2e']['value'],3), 'warm', round(d['warm_ms_per_align'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
for r in 1 2; do
  RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_prev.so run prev
  run new
  RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_vB.so run kw_seeds32
  RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_vC.so run kw_seeds96
done 2>&1 | tee gpurun_out/r2c19_ab.txt
c4() { timeout 600 python tools/bench_c4.py --pairs 512 --batched 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['pairs_per_s'],1), d['recovered_truth'], {k: round(v,1) if isinstance(v,float) else v for k,v in d['stage_ms_rank0'].items()})"; }
for r in 1 2; do
  RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_prev.so c4 c4_prev
  c4 c4_new
done 2>&1 | tee gpurun_out/r2c19_c4.txt
