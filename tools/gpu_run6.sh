mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
echo "== default"; timeout 300 python tools/phase_probe.py 20000 2>&1 | tail -1 | tee gpurun_out/phase_probe.txt
echo "== cas_first"; ROGTK_CAS_FIRST=1 timeout 300 python tools/phase_probe.py 20000 2>&1 | tail -1 | tee -a gpurun_out/phase_probe.txt
ROGTK_CAS_FIRST=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
