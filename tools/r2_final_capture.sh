# round 2, final evidence: the two bench arms, launch list, ncu --set full of the shipped kernels (step / eager / large),
# device timeline, sanitizer runs.  Summaries are made on the box; .ncu-rep files are dropped if gpurun_out would exceed
# the 64 MiB return limit.
mkdir -p gpurun_out
P=gpurun_out/r2f
timeout 1200 python bench.py > ${P}_bench_default.json 2> ${P}_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_reference.json 2> ${P}_bench_reference.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches.csv python bench.py --steps 2 --warmup 2 --no-cpu --no-large --no-extra --concurrent 0 > ${P}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_knn_warp|k_correspond|k_trial_step|k_linearize|k_compute_error|k_covariance|k_rs_onesweep|k_keys_hist|k_build_tables|k_count_cells" -c 48 -o ${P}_prof_step python tools/prof_step.py 1 > ${P}_prof_step.log 2>&1
python tools/ncu_summary.py ${P}_prof_step.ncu-rep ${P}_step_kernels.txt k_trial_step 0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_knn_tile|k_knn_warp|k_covariance" -c 6 -o ${P}_prof_eager python tools/prof_step.py 1 eager > ${P}_prof_eager.log 2>&1
python tools/ncu_summary.py ${P}_prof_eager.ncu-rep ${P}_eager_kernels.txt k_knn_tile 0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_covariance|k_linearize|k_compute_error|k_correspond" -c 8 -o ${P}_prof_large python tools/roofline_large.py 16 > ${P}_prof_large.log 2>&1
python tools/ncu_summary.py ${P}_prof_large.ncu-rep ${P}_large_kernels.txt k_linearize 0
RGC_TIMELINE=1 timeout 300 python tools/prof_step.py 6 > ${P}_timeline.log 2> ${P}_timeline.err
python - <<'PY'
lines = open("gpurun_out/r2f_timeline.err").read().split("\n")
starts = [i for i, l in enumerate(lines) if "marks (us since" in l]
open("gpurun_out/r2f_timeline_cold_step.txt", "w").write("\n".join(lines[starts[-1]:]) + "\n")
PY
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > ${P}_sanitize_memcheck.log 2>&1; tail -2 ${P}_sanitize_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py > ${P}_sanitize_racecheck.log 2>&1; tail -2 ${P}_sanitize_racecheck.log
du -sm gpurun_out
for f in prof_step prof_eager prof_large; do
  if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f ${P}_$f.ncu-rep; echo "dropped $f.ncu-rep"; fi
done
tail -c 300 ${P}_bench_default.json
