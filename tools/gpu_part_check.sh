timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "partition or spectrum or cfg3 or skew" 2>&1 | tail -3
bash tools/gpu_cfg3_launches.sh 2>&1 | head -6
python tools/big_probe.py 30000 | tail -1 | cut -c1-200
