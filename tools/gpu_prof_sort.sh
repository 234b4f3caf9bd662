#!/bin/bash
# ncu --set full of k_sort_small (the second kernel of the step) on 20 k UMIs of the headline shape
mkdir -p gpurun_out
CMD="python bench.py --umis 20000 --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout 300 $CMD > gpurun_out/plain_sort.log 2>&1 || { tail -5 gpurun_out/plain_sort.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sort_small -s 1 -c 1 -o gpurun_out/prof_sort_r01k $CMD > gpurun_out/ncu_sort.log 2>&1
tail -1 gpurun_out/ncu_sort.log
