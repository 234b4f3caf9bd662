#!/bin/bash
# parity suite + 200 k-UMI device bench + read-length sweep (the inner loop of hot-kernel work)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
timeout 300 python bench.py --umis 200000 --no-cpu --no-e2e --steps 5 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f G obs/s  ms %.2f  kernel_ms %.2f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['kernel_ms']))"
timeout 300 python tools/len_sweep.py 100000 2>&1 | tail -5
