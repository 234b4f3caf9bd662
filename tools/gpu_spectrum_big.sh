#!/bin/bash
# usage: gpu_spectrum_big.sh N TEMPLATES_PER_RANK  -- device-resident global spectrum at a fraction of BASELINE configs[4]: TEMPLATES_PER_RANK x N
# templates, 50 / N reads of each per rank, rounds of 1.5 M reads, 2 iterations (the second is reported)
N=${1:-8}; T=${2:-600000}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 tools/spectrum_ngpu.py $((T * N)) 1500000 0 2 2>&1 | grep -v "^\[" | tail -2 | tee gpurun_out/spectrum_big_n$N.json
