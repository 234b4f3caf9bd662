#!/bin/bash
# ncu --set full of the fused per-group unitig kernel (50 k UMIs of the headline shape)
mkdir -p gpurun_out
CMD="python tools/unitig_probe.py 50000"
timeout 300 $CMD > gpurun_out/unitig_plain.log 2>&1 || { tail -5 gpurun_out/unitig_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_unitig_group -s 2 -c 1 -o gpurun_out/prof_unitig $CMD > gpurun_out/ncu_unitig.log 2>&1
tail -3 gpurun_out/unitig_plain.log; tail -1 gpurun_out/ncu_unitig.log
