#!/bin/bash
# gpurun with retries while the pod answers "transient" (nothing charged): tools/gpurun_retry.sh <log> <timeout> <command...>
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1
  if grep -q "status=transient" $log; then sleep 150; continue; fi
  break
done
