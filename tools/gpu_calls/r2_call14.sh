# round 2, GPU call 14: device timeline of a cold C2 step (RGC_TIMELINE=1), flushed and unflushed L2
mkdir -p gpurun_out
RGC_TIMELINE=1 timeout 300 python tools/prof_step.py 6 > gpurun_out/r2c14_timeline.log 2> gpurun_out/r2c14_timeline.err
python - <<'PY'
lines = open("gpurun_out/r2c14_timeline.err").read().split("\n")
starts = [i for i, l in enumerate(lines) if "marks (us since" in l]
print("\n".join(lines[starts[-1]:]))
PY
