# round 2 (third session), call 3 (2 GPUs): the driver's N>1 launch of both arms, and the sharded test script
N=2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu/run_sharded.py > gpurun_out/r3c3_sharded_2gpu.json 2> gpurun_out/r3c3_sharded_2gpu.err; echo "sharded rc=$?"; tail -c 600 gpurun_out/r3c3_sharded_2gpu.json
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 16 > gpurun_out/r3c3_bench_${N}gpu.json 2> gpurun_out/r3c3_bench_${N}gpu.err; echo "bench$N rc=$? wall ${SECONDS}s"
grep "bench +" gpurun_out/r3c3_bench_${N}gpu.err | tail -12
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r3c3_bench_${N}gpu.json").read().strip().split("\n")[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print("c4", d.get("c4")); print("c5", d.get("c5"))
except Exception as e:
    print("parse failed", e)
PY
SECONDS=0
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r3c3_ref_${N}gpu.json 2> gpurun_out/r3c3_ref_${N}gpu.err; echo "ref$N rc=$? wall ${SECONDS}s"
tail -c 300 gpurun_out/r3c3_ref_${N}gpu.json
