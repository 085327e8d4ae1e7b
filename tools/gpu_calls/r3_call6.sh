# round 2 (third session), call 6 (8 GPUs): flag-in-data peer-memory all-reduce at N=8: sharded parity script, then the bench (C4 ramped chunks, C5 50M)
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu/run_sharded.py > gpurun_out/r3c6_sharded_${N}gpu.json 2> gpurun_out/r3c6_sharded_${N}gpu.err; echo "sharded rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r3c6_sharded_${N}gpu.json").read().strip().split("\n")[-1])
    print({k: d[k] for k in d if k in ("allreduce","allreduce_us","ranks_agree","lin_H_rel","pose_dt","oracle_pose_dt","oracle_iters_equal","iters_equal","repeat_identical","c4_identical_to_single_rank","sharded_align_s","unsharded_align_s")})
except Exception as e:
    print("parse failed", e)
PY
SECONDS=0
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 16 > gpurun_out/r3c6_bench_${N}gpu.json 2> gpurun_out/r3c6_bench_${N}gpu.err; echo "bench$N rc=$? wall ${SECONDS}s"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r3c6_bench_${N}gpu.json").read().strip().split("\n")[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print("c4", d.get("c4")); print("c5", d.get("c5"))
except Exception as e:
    print("parse failed", e)
PY
