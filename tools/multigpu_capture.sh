# One 8-GPU session: scaling of the independent-registration configs and the sharded-map checks.
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  if [ $n = 1 ]; then timeout 300 python bench.py --gpus 1 --steps 20 --warmup 4 --no-cpu --concurrent 0 2>/dev/null | tail -1 > gpurun_out/scale_$n.json
  else timeout 300 $TR --nproc-per-node $n --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 4 --no-cpu --concurrent 0 2>/dev/null | tail -1 > gpurun_out/scale_$n.json; fi
done
for n in 1 8; do
  if [ $n = 1 ]; then timeout 300 python tools/bench_c4.py --pairs 512 --threads 4 2>/dev/null | tail -1 > gpurun_out/c4_$n.json
  else timeout 300 $TR --nproc-per-node $n --master-port 29610 tools/bench_c4.py --pairs 4096 --threads 4 2>/dev/null | tail -1 > gpurun_out/c4_$n.json; fi
done
timeout 300 $TR --nproc-per-node 2 --master-port 29620 tests/multigpu/run_sharded.py > gpurun_out/sharded2.log 2>&1
tail -n 5 gpurun_out/sharded2.log
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        d = json.load(open(f"gpurun_out/scale_{n}.json")); print(n, round(d["value"], 1), round(d["ms_per_step"], 3), round(d["e2e"]["value"], 1))
    except Exception as e: print(n, "failed", e)
for n in (1, 8):
    try: print("c4", n, open(f"gpurun_out/c4_{n}.json").read().strip()[:400])
    except Exception as e: print("c4", n, "failed", e)
PY
