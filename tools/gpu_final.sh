#!/bin/bash
# round-end evidence, part 1: full bench (both arms), phase cycles, other configs, ncu launch list.   usage: gpu_final.sh TAG
mkdir -p gpurun_out
TAG=$1
timeout 1200 python bench.py > gpurun_out/bench_full_$TAG.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null
timeout 300 python tools/phase_probe.py 20000 2>&1 | tail -1 > gpurun_out/phase_$TAG.txt
timeout 600 python tools/run_configs.py 4000 > gpurun_out/configs_$TAG.json 2> gpurun_out/configs.err; tail -2 gpurun_out/configs.err
CMD="python bench.py --umis 100000 --steps 2 --warmup 1 --no-e2e --no-cpu"
timeout 300 $CMD > gpurun_out/plain_l.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l.log 2>&1
python -c "
import json; d=json.load(open('gpurun_out/bench_full_$TAG.json'))
print('value', d['value']/1e9, 'G obs/s  ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e ms', d['e2e']['ms_per_step'], 'e2e', d['e2e']['value']/1e9, 'cpu', d['cpu_baseline']['value']/1e9, d['cpu_baseline']['cores'])"
