timeout 300 python tools/phase_probe.py 20000 2>&1 | tail -1
