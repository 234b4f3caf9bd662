#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout_s> '<command>'   -- retries while the pod answers busy (exit 3 / transient), logs to <log>
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if ! grep -q "status=transient\|retry in a few minutes\|no box or slot" $log; then exit 0; fi
  sleep 45
done
