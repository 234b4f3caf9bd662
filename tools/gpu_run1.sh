mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
free -g | head -2; nproc
./tools/microbench_smem > gpurun_out/microbench_smem.txt 2>&1; cat gpurun_out/microbench_smem.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --umis 100000 --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_100k.json 2> gpurun_out/bench_100k.err; tail -3 gpurun_out/bench_100k.err; cat gpurun_out/bench_100k.json
