# A/B of two builds inside one GPU session: tools/ab.sh <rounds>; prev = rgc_slam_b200/librgc_gicp_prev.so
run() { timeout 300 python bench.py --steps 30 --warmup 4 --no-cpu --concurrent 0 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), round(1e3/d['e2e']['value'],3), round(d['warm_ms_per_align'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
for r in $(seq 1 ${1:-2}); do
  RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_prev.so run prev
  run new

done
