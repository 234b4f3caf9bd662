#!/bin/bash
# inner loop of the round-2 hot-kernel work: parity suite with the link-walk kernel, then 200 k-UMI device bench per range length
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
for rl in 0 ${RLS:-3 4}; do
  echo "== ROGTK_WALK_RL=$rl"
  ROGTK_WALK_RL=$rl timeout 300 python bench.py --umis 200000 --no-cpu --no-e2e --steps 5 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f G obs/s  ms %.2f  kernel_ms %.2f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['kernel_ms']))"
done
for rl in ${RLS:-3 4}; do ROGTK_WALK_RL=$rl timeout 300 python tools/walk_stats.py 20000 2>&1 | tail -2; done
ROGTK_WALK_RL=0 timeout 300 python tools/k_sweep.py 100000 2>&1 | tail -3
ROGTK_WALK_RL=4 timeout 300 python tools/k_sweep.py 100000 2>&1 | tail -3
