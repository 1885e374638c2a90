#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> <log> '<command>'   -- retries gpurun while the pod answers "busy / draining"
T=$1; LOG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > "$LOG" 2>&1
  if grep -q "status=transient\|status=busy\|status=refused\|nothing was charged" "$LOG"; then sleep 45; continue; fi
  break
done
tail -5 "$LOG"
