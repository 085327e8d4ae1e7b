# round 2, GPU call 10: evidence for profiles/ -- default bench line, launch list, ncu --set full of the shipped kernels
# (summaries are made on the box; the .ncu-rep files are kept only while gpurun_out stays under the 64 MiB return limit)
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r2c10_bench_default.json 2> gpurun_out/r2c10_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c10_bench_reference.json 2> gpurun_out/r2c10_bench_reference.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c10_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-large --no-extra --concurrent 0 > gpurun_out/r2c10_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_knn_tile|k_knn_warp|k_correspond|k_linearize|k_compute_error|k_covariance" -c 30 -o gpurun_out/r2c10_prof_step python tools/prof_step.py 1 > gpurun_out/r2c10_prof_step.log 2>&1
python tools/ncu_summary.py gpurun_out/r2c10_prof_step.ncu-rep gpurun_out/r2c10_step_kernels.txt k_correspond 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_knn_tile|k_knn_warp|k_covariance" -c 6 -o gpurun_out/r2c10_prof_eager python tools/prof_step.py 1 eager > gpurun_out/r2c10_prof_eager.log 2>&1
python tools/ncu_summary.py gpurun_out/r2c10_prof_eager.ncu-rep gpurun_out/r2c10_eager_kernels.txt k_knn_tile 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_covariance|k_linearize|k_compute_error|k_correspond" -c 8 -o gpurun_out/r2c10_prof_large python tools/roofline_large.py 16 > gpurun_out/r2c10_prof_large.log 2>&1
python tools/ncu_summary.py gpurun_out/r2c10_prof_large.ncu-rep gpurun_out/r2c10_large_kernels.txt k_linearize 0
ls -la gpurun_out/
du -sm gpurun_out
# keep the total under the return limit: drop the largest reports first
for f in r2c10_prof_step r2c10_prof_eager r2c10_prof_large; do
  if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/$f.ncu-rep; echo "dropped $f.ncu-rep"; fi
done
tail -c 400 gpurun_out/r2c10_bench_default.json
