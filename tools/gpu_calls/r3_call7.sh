# round 2 (third session), call 7: final state on one GPU — the GPU suite, the default bench line, the ncu launch list
mkdir -p gpurun_out
timeout 85 python -m pytest tests -m gpu -x -q > gpurun_out/r3c7_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r3c7_pytest.log
tail -4 gpurun_out/r3c7_pytest.log
timeout 75 python bench.py > gpurun_out/r3c7_bench_default.json 2> gpurun_out/r3c7_bench_default.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r3c7_bench_default.json
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3c7_launches.csv python bench.py --steps 2 --warmup 2 --no-cpu --no-large --no-extra --concurrent 0 > gpurun_out/r3c7_bench_under_ncu.log 2>&1; echo "ncu rc=$?"
