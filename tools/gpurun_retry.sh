#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers "transient" (nothing charged)
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=transient" "$LOG"; then sleep 45; continue; fi
  break
done
