#!/bin/bash
# GPU parity suite, the multi-pass probe and the non-headline configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
timeout 600 python tools/multipass_probe.py 2>&1 | tail -8 | tee gpurun_out/multipass_probe.txt
timeout 600 python tools/run_configs.py 4000 > gpurun_out/configs.json 2> gpurun_out/configs.err
python -c "
import json
for l in open('gpurun_out/configs.json'):
    d=json.loads(l); print(d['config'], d['k'], 'spilled', d['spilled_groups'], 'device_ms %.2f' % d['device_ms'], '%.2f Gobs/s' % (d['obs_per_s_device']/1e9), d['parity_first_groups'])"
