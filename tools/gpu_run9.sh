mkdir -p gpurun_out
timeout 1500 python tools/run_configs.py 4000 2>&1 | tail -12
