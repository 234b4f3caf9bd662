#!/bin/bash
# round-end evidence, part 2: ncu --set full of the dominant kernel (after the same command ran clean).   usage: gpu_final_ncu.sh TAG
mkdir -p gpurun_out
TAG=$1
CMD2="python bench.py --umis 20000 --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout 300 $CMD2 > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_smem_count -s 1 -c 1 -o gpurun_out/prof_$TAG $CMD2 > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/ncu.log
