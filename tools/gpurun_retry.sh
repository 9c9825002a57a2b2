#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <log-file> <command...>   -- retries while the pod has no free slot
t=$1; log=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@" > "$log" 2>&1
  if grep -q "status=transient\|status=busy\|rc=3" "$log" && ! grep -q "status=ok" "$log"; then sleep 90; continue; fi
  break
done
