#!/bin/bash
# like tools/gpu.sh, but retries while the pod answers busy (nothing is charged for those):  tools/gpu_retry.sh [--gpus N] TIMEOUT 'command'
cd "$(dirname "$0")/.."
for i in $(seq 1 40); do
  out=$(tools/gpu.sh "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|backing off"; then sleep 120; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
