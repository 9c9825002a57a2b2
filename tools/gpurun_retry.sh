#!/bin/bash
# usage: [GPURUN_OPTS="--gpus 2"] tools/gpurun_retry.sh <timeout-seconds> <log-file> <command>   -- retries while the pod has no free slot
t=$1; log=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GPURUN_OPTS --timeout "$t" -- "$@" > "$log" 2>&1
  if grep -q "status=transient\|status=busy" "$log" && ! grep -q "status=ok\|status=fail" "$log"; then sleep 90; continue; fi
  break
done
