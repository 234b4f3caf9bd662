mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
( time timeout 1200 python bench.py --no-cpu > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | tail -4
tail -5 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
