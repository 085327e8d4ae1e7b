# re-entry check: GPU suite, smoke, both bench arms on the restored HEAD
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3c1_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r3c1_pytest.log
tail -4 gpurun_out/r3c1_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3c1_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r3c1_smoke.log
timeout 900 python bench.py > gpurun_out/r3c1_bench_default.json 2> gpurun_out/r3c1_bench_default.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r3c1_bench_default.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3c1_bench_reference.json 2> gpurun_out/r3c1_bench_reference.err; echo "ref rc=$?"
