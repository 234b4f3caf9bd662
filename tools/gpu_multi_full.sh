mkdir -p gpurun_out
N=$1
free -g | head -2
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_full_n$N.json 2> gpurun_out/bench_full_n$N.err ) 2>&1 | tail -3
echo rc=$?; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_full_n$N.err | tail -5; cat gpurun_out/bench_full_n$N.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N',d['n_gpus'],'value %.1f G obs/s'%(d['value']/1e9),'ms/step %.1f'%d['ms_per_step'],'e2e %.1f G obs/s'%(d['e2e']['value']/1e9),'e2e umis/rank',d['e2e']['umis_per_rank'],'e2e ms %.1f'%d['e2e']['ms_per_step'], 'clocks', d['clocks'])"
