#!/bin/bash
# usage: gpu_spectrum.sh N  -- the global-spectrum mode on N GPUs: parity runs (reads from the host, checked against the oracle, one round and
# 20 rounds), then a device-resident timing run: 10 M reads per rank (200 000 x N templates, 50 / N reads of each per rank) in rounds of 1.5 M
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "spectrum" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/spectrum_ngpu.py 2000 0 1 2>&1 | grep -v "^\[" | tail -1 | tee gpurun_out/spectrum_parity_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/spectrum_ngpu.py 20000 25000 1 2>&1 | grep -v "^\[" | tail -1 | tee -a gpurun_out/spectrum_parity_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 tools/spectrum_ngpu.py $((200000 * N)) 1500000 0 2>&1 | grep -v "^\[" | tail -1 | tee gpurun_out/spectrum_timing_n$N.json
