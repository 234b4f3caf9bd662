mkdir -p gpurun_out
( time timeout 1200 python bench.py --no-cpu > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | tail -4
tail -3 gpurun_out/bench_full.err; python -c "
import json; d=json.load(open('gpurun_out/bench_full.json'))
print('value', d['value']/1e9, 'G obs/s  ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e ms', d['e2e']['ms_per_step'], 'e2e G', d['e2e']['value']/1e9)"
CMD="python bench.py --umis 100000 --steps 2 --warmup 1 --no-e2e --no-cpu"
timeout 300 $CMD > gpurun_out/plain_l.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01e.csv $CMD > gpurun_out/ncu_l.log 2>&1
