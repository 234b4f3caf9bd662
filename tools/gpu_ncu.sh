#!/bin/bash
# ncu --set full captures of the step's kernels on a 100 k-UMI launch (after the same command ran clean without ncu)
mkdir -p gpurun_out
CMD="python tools/csr_probe.py 100000"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 || exit 1
for spec in "smem:k_smem_count" "sort:k_sort_small" "csr:k_csr_group"; do
  tag=${spec%%:*}; kn=${spec##*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kn -s 2 -c 1 -f -o gpurun_out/r02_$tag $CMD > gpurun_out/ncu_$tag.log 2>&1
  ls -la gpurun_out/r02_$tag.ncu-rep
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_100k.csv python bench.py --umis 100000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_l.log 2>&1
