#!/bin/bash
# usage: gpu_spectrum.sh N  -- the global-spectrum mode on N GPUs: parity run (2000 templates), then a timing run
N=${1:-2}
timeout 600 python -m pytest tests -m gpu -x -q -k "spectrum" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/spectrum_ngpu.py 2000 0 1 2>&1 | grep -v "^\[" | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/spectrum_ngpu.py 20000 25000 1 2>&1 | grep -v "^\[" | tail -3
