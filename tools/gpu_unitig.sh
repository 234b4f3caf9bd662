#!/bin/bash
# GPU parity suite + unitig probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
timeout 600 python tools/unitig_probe.py 200000 2>&1 | tail -3 | tee gpurun_out/unitig_probe.txt
