# One GPU session producing the artefacts copied into profiles/ (see profiles/README.md)
set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu --concurrent 0 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_knn_tile|k_knn_warp|k_correspond|k_linearize|k_compute_error|k_covariance" -c 40 -o gpurun_out/prof_final python tools/prof_step.py 1 > gpurun_out/prof_final.log 2>&1
timeout 300 python tools/roofline_large.py 16 > gpurun_out/roofline_large.json 2> gpurun_out/roofline_large.err
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/sanitize_racecheck.log 2>&1
tail -3 gpurun_out/sanitize_memcheck.log gpurun_out/sanitize_racecheck.log
tail -c 600 gpurun_out/bench_default.json
