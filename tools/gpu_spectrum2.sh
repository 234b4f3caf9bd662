#!/bin/bash
# global k-mer spectrum on 2 GPUs over NCCL: one round, then chunked rounds (share of the spectrum carried in HBM)
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/spectrum_2gpu.py "$@" 2>gpurun_out/spec.err | grep global_spectrum; }
run 2000 > gpurun_out/spectrum_2gpu.txt
run 2000 20000 >> gpurun_out/spectrum_2gpu.txt
run 20000 100000 >> gpurun_out/spectrum_2gpu.txt
cat gpurun_out/spectrum_2gpu.txt; tail -3 gpurun_out/spec.err
