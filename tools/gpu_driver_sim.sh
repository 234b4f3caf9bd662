#!/bin/bash
# what the driver runs at round end, in its order: smoke(), pytest -m gpu, reference arm, our arm (defaults)
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 900 python bench.py --impl reference > gpurun_out/sim_ref.json 2>/dev/null; echo "ref rc=$?"; cut -c1-200 gpurun_out/sim_ref.json
timeout 900 python bench.py > gpurun_out/sim_ours.json 2>/dev/null; echo "ours rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/sim_ours.json')); print(len(open('gpurun_out/sim_ours.json').read().strip().splitlines()), 'line(s);', 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'frac %.4f' % d['roofline']['frac'], 'issue %.3f' % d['roofline']['issue']['frac'], 'launches', d['gpu_launches'], d['clocks']['reasons'], 'cpu %.3g' % d['cpu_baseline']['value'])"
