# round 2, GPU call 4: parity suite + launch timeline of one batched chunk
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2c4_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c4_pytest.log
tail -6 gpurun_out/r2c4_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2c4_batch_launches.csv python tools/bench_c4.py --pairs 64 --batched > gpurun_out/r2c4_batch_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2c4_batch_launches.csv")) if len(r)>5 and r[0].isdigit()]
names=[(r[4].split("(")[0].replace("rgc::","").replace("void ",""), float(r[-1]), r[8]) for r in rows]
# the timed call is the last rgc_batch_align: starts at the last-but-one k_bingest pair
idx=[i for i,(n,_,_) in enumerate(names) if n=="k_bingest"]
start=idx[-2]
tot=collections.OrderedDict()
for n,t,g in names[start:]:
    tot[n]=tot.get(n,0)+t
print("launches in the timed call:", len(names)-start)
for n,t in sorted(tot.items(), key=lambda x:-x[1]): print(f"{n:28s} {t/1e3:10.1f} us")
print("--- LM rounds in order")
for n,t,g in names[start:]:
    if n in ("k_bcorrespond","k_knn_warp","k_blinearize","k_bcompute_error","k_bfitness","k_knn_tile") or n.startswith("k_cov"):
        print(f"{n:22s} {g:16s} {t/1e3:9.1f} us")
PY
