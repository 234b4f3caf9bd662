mkdir -p gpurun_out
timeout 600 python tools/cfg3_once.py > gpurun_out/plain_cfg3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spill_extract -s 2 -c 1 -o gpurun_out/prof_extract python tools/cfg3_once.py > gpurun_out/ncu_x.log 2>&1
tail -2 gpurun_out/ncu_x.log
