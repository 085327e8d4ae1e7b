# register-cap / gather-batch sweep of k_covariance at 8M points (builds each variant on the box)
run() { timeout 300 python tools/roofline_large.py 16 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', {k: (round(d[k]['ms'],3), round(d[k].get('frac_of_peak',0),3)) for k in ('k_covariance','k_linearize','k_compute_error','k_knn_tile')})"; }
for v in "-DRGC_COV_MINB=6 -DRGC_COV_BATCH=5" "-DRGC_COV_MINB=5 -DRGC_COV_BATCH=5" "-DRGC_COV_MINB=4 -DRGC_COV_BATCH=5" "-DRGC_COV_MINB=5 -DRGC_COV_BATCH=10"; do
  RGC_NVCC_EXTRA="$v" python -m rgc_slam_b200.build > /dev/null 2>&1 && run "$v"
done
python -m rgc_slam_b200.build > /dev/null 2>&1
