#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout_s> '<command>'   -- retries while the pod answers "transient"/busy
log=$1; to=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  if grep -q "status=transient\|status=busy\|no box or slot" "$log"; then sleep 90; continue; fi
  break
done
echo "[retry] done after $i attempt(s)" >> "$log"
