#!/bin/bash
# round-2 evidence on one GPU: full default bench (both arms), then ncu launch list + --set full captures of the step's kernels
bash tools/gpu_bench_full.sh r02
bash tools/gpu_ncu.sh
