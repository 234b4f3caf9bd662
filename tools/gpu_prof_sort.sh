#!/bin/bash
# launch list of a min_coverage=1 step (mid-size counting sort tier) on 50 k UMIs
mkdir -p gpurun_out
CMD="python tools/unitig_probe.py 50000"
timeout 300 $CMD > gpurun_out/sort_plain.log 2>&1 || { tail -5 gpurun_out/sort_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_sortmid.csv $CMD > gpurun_out/ncu_sortmid.log 2>&1
tail -3 gpurun_out/sort_plain.log
