#!/bin/bash
# launch list (per-kernel times) of the global-spectrum local count on one giant group
mkdir -p gpurun_out
CMD="python tools/spectrum_probe.py 4000"
timeout 300 $CMD > gpurun_out/spill_plain.log 2>&1 || { tail -5 gpurun_out/spill_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_spectrum.csv $CMD > gpurun_out/ncu_spill.log 2>&1
tail -2 gpurun_out/spill_plain.log
