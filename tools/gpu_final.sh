# round-end evidence: full bench (both arms), launch list, full ncu capture of the dominant kernel, phase cycles
mkdir -p gpurun_out
TAG=$1
timeout 1200 python bench.py > gpurun_out/bench_full_$TAG.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null
timeout 300 python tools/phase_probe.py 20000 2>&1 | tail -1 > gpurun_out/phase_$TAG.txt
CMD="python bench.py --umis 100000 --steps 2 --warmup 1 --no-e2e --no-cpu"
timeout 300 $CMD > gpurun_out/plain_l.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l.log 2>&1
CMD2="python bench.py --umis 20000 --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout 300 $CMD2 > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_smem_count -s 1 -c 1 -o gpurun_out/prof_$TAG $CMD2 > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/ncu.log
python -c "
import json; d=json.load(open('gpurun_out/bench_full_$TAG.json'))
print('value', d['value']/1e9, 'G obs/s  ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e ms', d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value']/1e9, d['cpu_baseline']['cores'])"
