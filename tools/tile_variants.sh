run() { timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms'].items() if 'knn' in k})"; }
B="-DRGC_KT_PEND=8 -DRGC_KT_STACK=40"
for v in "$B -DRGC_KT_CAND=96 -DRGC_KT_LEAF=64" "$B -DRGC_KT_CAND=128 -DRGC_KT_LEAF=96" "$B -DRGC_KT_CAND=128 -DRGC_KT_LEAF=48 -DRGC_KT_SEEDS=96" "-DRGC_KT_PEND=4 -DRGC_KT_STACK=40 -DRGC_KT_CAND=96 -DRGC_KT_LEAF=48" "$B -DRGC_KT_CAND=128 -DRGC_KT_LEAF=64 -DRGC_KT_SEEDS=128"; do
  RGC_NVCC_EXTRA="$v" python -m rgc_slam_b200.build > /dev/null 2>&1 && run "$v"
done
