#!/bin/bash
# launch list of the unitig-compaction kernels (50 k UMIs of the headline shape)
mkdir -p gpurun_out
CMD="python tools/unitig_probe.py 50000"
timeout 300 $CMD > gpurun_out/unitig_plain.log 2>&1 || { tail -5 gpurun_out/unitig_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_unitig.csv $CMD > gpurun_out/ncu_unitig.log 2>&1
tail -3 gpurun_out/unitig_plain.log
