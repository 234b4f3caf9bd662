mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --umis 200000 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "stdout lines: $(wc -l < gpurun_out/bench_n2.json)"; head -c 300 gpurun_out/bench_n2.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
echo "ref stdout lines: $(wc -l < gpurun_out/bench_ref_n2.json)"; head -c 400 gpurun_out/bench_ref_n2.json; echo
