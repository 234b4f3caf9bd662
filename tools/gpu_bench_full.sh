mkdir -p gpurun_out
set -x
free -g | head -2
( time timeout 1200 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | tail -4
tail -5 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | tail -4
cat gpurun_out/bench_ref.json
