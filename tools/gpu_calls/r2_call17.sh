# round 2, GPU call 17: group seeds for the unhinted search; float->double widening on the integer pipe (variant vC) on the HBM kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c17_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c17_pytest.log
tail -6 gpurun_out/r2c17_pytest.log
run() { timeout 300 python bench.py --steps 60 --warmup 8 --no-cpu --no-large --no-extra --concurrent 0 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), 'p50', round(d['p50_ms'],3), 'e2e ms', round(1e3/d['e2e']['value'],3), 'warm', round(d['warm_ms_per_align'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
for r in 1 2; do
  RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_prev.so run prev
  run new
done 2>&1 | tee gpurun_out/r2c17_ab.txt
large() { timeout 300 python tools/roofline_large.py 16 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', {k: (round(d[k]['ms'],4), round(d[k].get('frac_of_peak',0),3)) for k in ('k_covariance','k_linearize','k_compute_error','k_correspond','k_knn_tile')})"; }
for r in 1 2; do
  large default
  RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_vC.so large f2d_bits
done 2>&1 | tee gpurun_out/r2c17_large.txt
