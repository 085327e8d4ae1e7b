# round 2, GPU call 16: group-cooperative 1-NN (4 / 8 lanes per query), mapped result slots, feature clouds (less-flat / ground lists)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c16_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c16_pytest.log
tail -6 gpurun_out/r2c16_pytest.log
run() { timeout 300 python bench.py --steps 60 --warmup 8 --no-cpu --no-large --no-extra --concurrent 0 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), 'p50', round(d['p50_ms'],3), 'e2e ms', round(1e3/d['e2e']['value'],3), 'warm', round(d['warm_ms_per_align'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
for r in 1 2; do
  RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_prev.so run prev
  run new
  RGC_LIB=$PWD/rgc_slam_b200/librgc_gicp_vB.so run new_nogroup
done 2>&1 | tee gpurun_out/r2c16_ab.txt
RGC_TIMELINE=1 timeout 300 python tools/prof_step.py 6 > gpurun_out/r2c16_timeline.log 2> gpurun_out/r2c16_timeline.err
python - <<'PY'
lines = open("gpurun_out/r2c16_timeline.err").read().split("\n")
starts = [i for i, l in enumerate(lines) if "marks (us since" in l]
print("\n".join(lines[starts[-1]:]))
PY
