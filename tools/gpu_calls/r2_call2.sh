# round 2, GPU call 2: parity suite again + launch timeline of one cold align + source-kNN defer A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c2_pytest.log
tail -4 gpurun_out/r2c2_pytest.log
run() { tag=$1; shift; env "$@" timeout 400 python bench.py --steps 30 --warmup 4 --no-cpu --concurrent 0 2> gpurun_out/r2c2_$tag.err | tail -1 > gpurun_out/r2c2_$tag.json; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2c2_$tag.json"))
    print("$tag", "ms/step", round(d["ms_per_step"],3), "e2e ms", round(1e3/d["e2e"]["value"],3), "warm", round(d["warm_ms_per_align"],3), "vgicp", round(d["vgicp"]["cold_ms_per_align"],3), {k: round(v,3) for k,v in d["stage_ms"].items()}, "launches", d["gpu_launches"])
except Exception as e:
    print("$tag failed", e)
PY
}
run default A=1
run defer1 RGC_KNN_DEFER=1
run defer150 RGC_KNN_DEFER=150
run defer300 RGC_KNN_DEFER=300
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --concurrent 0 > gpurun_out/r2c2_bench_under_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2c2_launches.csv")) if len(r)>5 and r[0].isdigit()]
names=[(r[4].split("(")[0].replace("rgc::",""), float(r[-1])) for r in rows]
idx=[i for i,(n,_) in enumerate(names) if n.startswith("k_ingest")]
print("n launches", len(names))
start=idx[-2] if len(idx)>=2 else 0
tot=0
for n,t in names[start:start+140]:
    print(f"{n:28s} {t/1e3:8.1f} us"); tot+=t
print("sum", tot/1e3)
PY
