mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 300 python tools/phase_probe.py 20000 2>&1 | tee gpurun_out/phase_probe.txt
timeout 600 python bench.py --umis 100000 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_100k.json 2> gpurun_out/bench_100k.err; tail -3 gpurun_out/bench_100k.err; cat gpurun_out/bench_100k.json
