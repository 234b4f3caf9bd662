#!/bin/bash
# launch list of one skewed-config (cfg3 subset) call: where the time of the three counting tiers goes
mkdir -p gpurun_out
CMD="python tools/cfg3_once.py"
timeout 300 $CMD > gpurun_out/cfg3_plain.log 2>&1 || { tail -5 gpurun_out/cfg3_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg3.csv $CMD > gpurun_out/ncu_cfg3.log 2>&1
tail -2 gpurun_out/cfg3_plain.log
