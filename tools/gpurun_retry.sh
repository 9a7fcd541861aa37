#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE [gpurun args...] : retries while the pod answers "transient" (nothing charged)
log=$1; shift
for i in $(seq 1 40); do
  gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient" "$log"; then sleep 45; continue; fi
  break
done
