#!/bin/bash
# usage: gpu_scale.sh N [steps]  -- the driver's N-GPU launch of bench.py (weak scaling: 1 M UMIs per rank), the one-process multi-device entry
# test, then the global-spectrum mode on N GPUs (parity vs the oracle + device-resident timing).  Outputs under gpurun_out/.
N=${1:-2}; STEPS=${2:-3}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps $STEPS --warmup 3 \
   > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value %.3e' % d['value'], 'ms', round(d['ms_per_step'],2), 'e2e %.3e' % d['e2e']['value'], 'e2e ms', round(d['e2e']['ms_per_step'],1))
print('next_rows', json.dumps(d.get('next_rows'))[:1500])
PY
timeout 600 python -m pytest tests -m gpu -x -q -k "several_gpus" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/spectrum_ngpu.py 2000 0 1 2>&1 | grep -v "^\[" | tail -1 | tee gpurun_out/spectrum_parity_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/spectrum_ngpu.py 20000 25000 1 2>&1 | grep -v "^\[" | tail -1 | tee -a gpurun_out/spectrum_parity_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 tools/spectrum_ngpu.py $((200000 * N)) 1500000 0 2>&1 | grep -v "^\[" | tail -1 | tee gpurun_out/spectrum_timing_n$N.json
