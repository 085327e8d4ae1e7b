# parameter sweep of the tile kNN kernel on the C2 workload (each variant is built on the box)
run() { timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --concurrent 0 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms'].items() if 'knn' in k})"; }
for v in "-DRGC_KT_PEND=8" "-DRGC_KT_PEND=12" "-DRGC_KT_PEND=12 -DRGC_KT_CAND=96" "-DRGC_KT_PEND=16 -DRGC_KT_CAND=96 -DRGC_KT_STACK=32"; do
  RGC_NVCC_EXTRA="$v" python -m rgc_slam_b200.build > /dev/null 2>&1 && run "$v"
done
python -m rgc_slam_b200.build > /dev/null 2>&1
