#!/bin/bash
# here (no GPU): .ncu-rep files of tools/gpu_ncu.sh -> profiles/ncu_counters.json (keyed by the current build id) + text summaries + launch list
python tools/ncu_counters.py 100000 "k_smem_count<1,32,false>"=gpurun_out/r02_smem.ncu-rep "k_sort_small"=gpurun_out/r02_sort.ncu-rep "k_csr_group"=gpurun_out/r02_csr.ncu-rep > /dev/null
for t in smem sort csr; do
  (echo "# ncu --set full --clock-control none --import-source on, one launch of the kernel at 100 k UMIs of the headline shape (tools/gpu_ncu.sh); summary by tools/ncu_summary.py"; python tools/ncu_summary.py gpurun_out/r02_$t.ncu-rep) > profiles/r02_ncu_${t}_full.txt 2>&1
done
cp gpurun_out/r02_launches_100k.csv profiles/r02_launches_100k.csv
python -c "
import json, sys; sys.path.insert(0, '.')
from bench import kernel_build_id
d = json.load(open('profiles/ncu_counters.json')); print('counters build', d['build_id'], 'tree build', kernel_build_id())"
