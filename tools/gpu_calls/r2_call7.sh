# round 2, GPU call 7: parity of the pipelined linearize + A/B at 8 M / 2 M points + C4 with a full-size warm-up
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2c7_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c7_pytest.log
tail -4 gpurun_out/r2c7_pytest.log
V=$PWD/rgc_slam_b200/variants
for tag in default lin_mb2 lin_mb4; do
  if [ $tag = default ]; then unset RGC_LIB; else export RGC_LIB=$V/$tag.so; fi
  timeout 300 python tools/roofline_large.py 16 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$tag', {k: (round(d[k]['ms'],4), round(d[k].get('frac_of_peak',0),3)) for k in ('k_covariance','k_linearize','k_compute_error','k_correspond','k_knn_tile')})"
done
unset RGC_LIB
timeout 600 python tools/bench_c4.py --pairs 512 --batched > gpurun_out/r2c7_c4.json 2>gpurun_out/r2c7_c4.err; tail -c 700 gpurun_out/r2c7_c4.json
