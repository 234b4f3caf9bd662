#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n tools/pcie_multi.py 2>/dev/null | grep aggregate | tee -a gpurun_out/pcie_multi.jsonl
done
nproc; grep -c processor /proc/cpuinfo; free -g | head -2
