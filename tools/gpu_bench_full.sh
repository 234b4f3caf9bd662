#!/bin/bash
# full default bench (both arms), as the driver runs it.   usage: gpu_bench_full.sh TAG
mkdir -p gpurun_out
TAG=${1:-x}
( time timeout 1500 python bench.py > gpurun_out/bench_full_$TAG.json 2> gpurun_out/bench_full.err ) 2>&1 | tail -3; tail -2 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_full_$TAG.json').read().strip().splitlines()[-1])
print('value %.3f G obs/s  ms %.2f  kernel_ms %.2f frac %.4f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))
for k in d['roofline']['kernels']: print('  ', k['kernel'], k['ms'], k['frac'])
print('with_csr', d['value_with_csr'])
print('e2e', {k:v for k,v in d['e2e'].items() if k!='variants'})
for v in d['e2e'].get('variants', []): print('   ', v)
print('cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], d['cpu_baseline'] and d['cpu_baseline']['cores'])
print('next', d['next_rows'])
print('clocks', d['clocks'])
r=json.loads(open('gpurun_out/bench_ref_$TAG.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline']['cores'])
PY
