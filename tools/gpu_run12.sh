mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/spectrum_2gpu.py 2000 > gpurun_out/spectrum_2gpu.txt 2>&1
grep -n "Error\|error\|Traceback\|File \"/root\|global_spectrum" gpurun_out/spectrum_2gpu.txt | head -30
