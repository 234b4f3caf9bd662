#!/bin/bash
# per-kernel times of the multi-pass path (launch list only)
mkdir -p gpurun_out
timeout 300 python tools/multipass_probe.py 100 400 > gpurun_out/multi_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_multi.csv \
    python tools/multipass_probe.py 100 400 > gpurun_out/multi_ncu.log 2>&1
tail -3 gpurun_out/multi_plain.log
