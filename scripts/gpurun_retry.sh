#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <gpurun args...>   - retries while the pod answers "transient" (nothing charged)
log=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > $log 2>&1
  if grep -q "status=transient\|status=refused" $log; then sleep 120; continue; fi
  break
done
tail -12 $log | cut -c1-400
