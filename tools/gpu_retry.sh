#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> [--gpus N] -- '<command>' : retries gpurun while the pod answers busy (rc 3 / transient)
T=$1; shift
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1); rc=$?
  echo "$out" | tail -25
  if echo "$out" | grep -q "status=transient"; then sleep 120; continue; fi
  if [ $rc -eq 3 ]; then sleep 120; continue; fi
  exit $rc
done
exit 3
