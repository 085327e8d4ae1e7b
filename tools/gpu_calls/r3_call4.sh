# round 2 (third session), call 4 (8 GPUs): the driver's N=8 launch of our arm (C2 replicas + C4 strong scaling + C5 50M-point map)
N=${1:-8}
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 16 > gpurun_out/r3c4_bench_${N}gpu.json 2> gpurun_out/r3c4_bench_${N}gpu.err; echo "bench$N rc=$? wall ${SECONDS}s"
grep "bench +" gpurun_out/r3c4_bench_${N}gpu.err | grep "rank 0" | tail -8
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r3c4_bench_${N}gpu.json").read().strip().split("\n")[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print("c4", d.get("c4")); print("c5", d.get("c5"))
except Exception as e:
    print("parse failed", e)
PY
