# launch list + full capture of the dominant kernel for profiles/ (run each under ncu only after the plain run exits 0)
mkdir -p gpurun_out
TAG=$1
CMD="python bench.py --umis 100000 --steps 2 --warmup 1 --no-e2e --no-cpu"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
CMD2="python bench.py --umis 20000 --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout 300 $CMD2 > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_smem_count -s 1 -c 1 -o gpurun_out/prof_$TAG $CMD2 > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
