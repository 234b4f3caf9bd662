#!/bin/bash
# per-chunk timeline of the host-buffer path (ROGTK_TRACE=1), then the GPU parity suite
mkdir -p gpurun_out
ROGTK_TRACE=1 timeout 600 python bench.py --umis 200000 --no-cpu --steps 1 --warmup 3 --e2e-steps 1 > gpurun_out/trace.json 2> gpurun_out/trace.log
tail -13 gpurun_out/trace.log
cat gpurun_out/trace.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'])"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
