#!/bin/bash
log=$1; shift; to=$1; shift; g=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $g --timeout $to -- "$@" > $log 2>&1
  if ! grep -q "status=transient\|retry in a few minutes\|no box or slot\|busy" $log; then exit 0; fi
  sleep 60
done
