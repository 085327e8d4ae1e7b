# round 2, GPU call 9 (2 GPUs): sharded registration vs unsharded vs oracle; bench --gpus 2 (C2 replicas + C4 strong + C5)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu/run_sharded.py > gpurun_out/r2c9_sharded.json 2> gpurun_out/r2c9_sharded.err; echo "sharded rc=$?"; tail -c 2500 gpurun_out/r2c9_sharded.json; tail -5 gpurun_out/r2c9_sharded.err
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 16 > gpurun_out/r2c9_bench2.json 2> gpurun_out/r2c9_bench2.err; echo "bench2 rc=$? wall ${SECONDS}s"; tail -4 gpurun_out/r2c9_bench2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2c9_bench2.json").read().strip().split("\n")[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print("c4", d.get("c4")); print("c5", d.get("c5"))
except Exception as e:
    print("parse failed", e)
PY
