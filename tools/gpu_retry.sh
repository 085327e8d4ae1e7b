#!/bin/bash
# tools/gpu_retry.sh <timeout> <script>: run a GPU call, retrying while the pod answers "busy" (nothing is charged then)
for i in $(seq 1 15); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out" | tail -${3:-60}
  exit 0
done
echo "gpu_retry: still busy after 15 attempts"
