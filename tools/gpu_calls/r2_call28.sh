# round 2, GPU call 28: single-block one-launch build of sweep-sized clouds (k_small_build), with speculative table sizes
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c28_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c28_pytest.log
tail -5 gpurun_out/r2c28_pytest.log
run() { timeout 300 python bench.py --steps 60 --warmup 8 --no-cpu --no-large --no-extra --concurrent 0 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), 'p50', round(d['p50_ms'],3), 'e2e ms', round(1e3/d['e2e']['value'],3), 'warm', round(d['warm_ms_per_align'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
for r in 1 2 3; do
  RGC_NO_SMALL_BUILD=1 run spec_tables_only
  run small_build
done 2>&1 | tee gpurun_out/r2c28_ab.txt
RGC_TIMELINE=1 timeout 300 python tools/prof_step.py 6 > gpurun_out/r2c28_timeline.log 2> gpurun_out/r2c28_timeline.err
python - <<'PY'
lines = open("gpurun_out/r2c28_timeline.err").read().split("\n")
starts = [i for i, l in enumerate(lines) if "marks (us since" in l]
print("\n".join(lines[starts[-1]:starts[-1]+40]))
PY
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/r2c28_memcheck.log 2>&1; tail -2 gpurun_out/r2c28_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/r2c28_racecheck.log 2>&1; tail -2 gpurun_out/r2c28_racecheck.log
