# round 2: multi-GPU evidence on N GPUs of one box: bash tools/r2_multigpu.sh N   (run as: gpurun --gpus N -- 'bash tools/r2_multigpu.sh N')
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu/run_sharded.py > gpurun_out/r2f_sharded_2gpu.json 2> gpurun_out/r2f_sharded_2gpu.err; echo "sharded rc=$?"; tail -c 1500 gpurun_out/r2f_sharded_2gpu.json
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
fi
SECONDS=0
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 16 > gpurun_out/r2f_bench_${N}gpu.json 2> gpurun_out/r2f_bench_${N}gpu.err; echo "bench$N rc=$? wall ${SECONDS}s"; tail -4 gpurun_out/r2f_bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2f_bench_${N}gpu.json").read().strip().split("\n")[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print("c4", d.get("c4")); print("c5", d.get("c5"))
except Exception as e:
    print("parse failed", e)
PY
