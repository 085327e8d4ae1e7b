# round 2, GPU call 25: coarse-to-fine first correspondence search (k_correspond_coarse), RGC_COARSE_SHIFT sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gicp.py tests/test_gpu_round2.py tests/test_gpu_batch.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2c25_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c25_pytest.log
tail -6 gpurun_out/r2c25_pytest.log
run() { timeout 300 python bench.py --steps 60 --warmup 8 --no-cpu --no-large --no-extra --concurrent 0 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), 'p50', round(d['p50_ms'],3), 'e2e ms', round(1e3/d['e2e']['value'],3), 'warm', round(d['warm_ms_per_align'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
for r in 1 2; do
  for s in 0 2 3 4 5; do RGC_COARSE_SHIFT=$s run shift$s; done
done 2>&1 | tee gpurun_out/r2c25_ab.txt
c4() { timeout 600 python tools/bench_c4.py --pairs 512 --batched 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['pairs_per_s'],1), d['recovered_truth'], {k: round(v,1) if isinstance(v,float) else v for k,v in d['stage_ms_rank0'].items()})"; }
for s in 0 2 3 4; do RGC_COARSE_SHIFT=$s c4 c4_shift$s; done 2>&1 | tee gpurun_out/r2c25_c4.txt
RGC_TIMELINE=1 timeout 300 python tools/prof_step.py 6 > gpurun_out/r2c25_timeline.log 2> gpurun_out/r2c25_timeline.err
python - <<'PY'
lines = open("gpurun_out/r2c25_timeline.err").read().split("\n")
starts = [i for i, l in enumerate(lines) if "marks (us since" in l]
print("\n".join(lines[starts[-1]:]))
PY
