# round 2, GPU call 3: full parity suite (batch path included) + C4 batched vs thread farm + headline
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2c3_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c3_pytest.log
tail -6 gpurun_out/r2c3_pytest.log
timeout 600 python tools/bench_c4.py --pairs 256 --batched > gpurun_out/r2c3_c4_batched.json 2> gpurun_out/r2c3_c4_batched.err; tail -c 1200 gpurun_out/r2c3_c4_batched.json; tail -3 gpurun_out/r2c3_c4_batched.err
timeout 600 python tools/bench_c4.py --pairs 256 --batched --chunk 32 > gpurun_out/r2c3_c4_batched32.json 2>/dev/null; tail -c 600 gpurun_out/r2c3_c4_batched32.json
timeout 600 python tools/bench_c4.py --pairs 256 --batched --chunk 128 > gpurun_out/r2c3_c4_batched128.json 2>/dev/null; tail -c 600 gpurun_out/r2c3_c4_batched128.json
timeout 600 python tools/bench_c4.py --pairs 256 --threads 8 > gpurun_out/r2c3_c4_threads8.json 2>/dev/null; tail -c 500 gpurun_out/r2c3_c4_threads8.json
timeout 400 python bench.py --steps 30 --warmup 4 --no-cpu --concurrent 0 2>/dev/null | tail -1 > gpurun_out/r2c3_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r2c3_bench.json')); print('bench ms/step', d['ms_per_step'], 'e2e ms', 1e3/d['e2e']['value'], 'warm', d['warm_ms_per_align'], d['stage_ms'])"
