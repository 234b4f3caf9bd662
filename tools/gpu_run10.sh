mkdir -p gpurun_out
timeout 600 python tools/cfg3_once.py > gpurun_out/plain_cfg3.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg3.csv python tools/cfg3_once.py > gpurun_out/ncu_cfg3.log 2>&1
tail -2 gpurun_out/ncu_cfg3.log
