# round 2 (third session), call 2: programmatic dependent launch on the build / LM chains; speculative tables only for
# clouds of the previous cloud's size (the 4096-pair C4 leg did not finish); bench.py with progress log + watchdog
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3c2_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r3c2_pytest.log
tail -5 gpurun_out/r3c2_pytest.log
run() { timeout 300 python bench.py --steps 60 --warmup 8 --no-cpu --no-large --no-extra --concurrent 0 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), 'p50', round(d['p50_ms'],3), 'e2e ms', round(1e3/d['e2e']['value'],3), 'warm', round(d['warm_ms_per_align'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
for r in 1 2; do
  RGC_NO_PDL=1 run no_pdl
  run pdl
done 2>&1 | tee gpurun_out/r3c2_ab.txt
SECONDS=0
timeout 700 python bench.py > gpurun_out/r3c2_bench_default.json 2> gpurun_out/r3c2_bench_default.err; echo "bench rc=$? wall ${SECONDS}s"
grep "bench +" gpurun_out/r3c2_bench_default.err | tail -12
tail -c 600 gpurun_out/r3c2_bench_default.json
