#!/bin/bash
# tools/gpurun_retry.sh <timeout-seconds> '<command>': gpurun with retries while the pod answers busy (nothing is charged for those)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "$out"; echo "gave up after 40 tries"; exit 3
