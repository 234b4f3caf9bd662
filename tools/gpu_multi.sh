mkdir -p gpurun_out
N=$1; U=$2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --umis $U --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo rc=$?; tail -5 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
