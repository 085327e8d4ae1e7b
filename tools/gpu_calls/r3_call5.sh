# round 2 (third session), call 5 (2 GPUs): all-reduce over peer memory (k_peer_allreduce, CUDA-IPC mailboxes) vs ncclAllReduce;
# ramped chunk sizes in rgc_batch_align
N=2
mkdir -p gpurun_out
for mode in p2p nccl; do
  if [ $mode = nccl ]; then export RGC_NO_P2P=1; else unset RGC_NO_P2P; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu/run_sharded.py > gpurun_out/r3c5_sharded_2gpu_$mode.json 2> gpurun_out/r3c5_sharded_2gpu_$mode.err; echo "sharded $mode rc=$?"
  tail -3 gpurun_out/r3c5_sharded_2gpu_$mode.err | cut -c1-300
  python - <<PY
import json
d=json.loads(open("gpurun_out/r3c5_sharded_2gpu_$mode.json").read().strip().split("\n")[-1])
print({k: d[k] for k in d if k in ("allreduce","allreduce_us","ranks_agree","lin_H_rel","pose_dt","oracle_pose_dt","oracle_iters_equal","iters_equal","repeat_identical","c4_identical_to_single_rank","sharded_align_s","unsharded_align_s")})
PY
done
unset RGC_NO_P2P
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 16 > gpurun_out/r3c5_bench_${N}gpu.json 2> gpurun_out/r3c5_bench_${N}gpu.err; echo "bench$N rc=$? wall ${SECONDS}s"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r3c5_bench_${N}gpu.json").read().strip().split("\n")[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print("c4", d.get("c4")); print("c5", d.get("c5"))
except Exception as e:
    print("parse failed", e)
PY
timeout 300 python -m pytest tests/test_gpu_batch.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -3
