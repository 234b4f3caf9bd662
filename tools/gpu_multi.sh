#!/bin/bash
# GPU parity suite, then the multi-pass probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
timeout 600 python tools/multipass_probe.py 2>&1 | tail -8 | tee gpurun_out/multipass_probe.txt
