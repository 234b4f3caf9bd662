mkdir -p gpurun_out
CMD="python bench.py --umis 20000 --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_smem_count -s 1 -c 1 -o gpurun_out/$1 $CMD > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
