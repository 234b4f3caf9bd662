#!/bin/bash
# parity suite + 200 k-UMI device bench (the inner loop of kernel work)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --umis ${UMIS:-200000} --steps 3 --e2e-steps 1 > gpurun_out/bench_200k.json 2> gpurun_out/bench_200k.err; tail -3 gpurun_out/bench_200k.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_200k.json').read().strip().splitlines()[-1])
print('value %.2f G obs/s  ms %.2f  kernel_ms %.2f frac %.4f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))
for k in d['roofline']['kernels']: print('  ', k['kernel'], k['ms'], k['frac'])
print('with_csr', d['value_with_csr'])
print('e2e', {k:v for k,v in d['e2e'].items() if k!='variants'})
for v in d['e2e'].get('variants', []): print('   ', v)
print('cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
PY
