mkdir -p gpurun_out
CMD2="python bench.py --umis 20000 --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout 300 $CMD2 > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sort_small -s 1 -c 1 -o gpurun_out/prof_sort $CMD2 > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
