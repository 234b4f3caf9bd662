mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
timeout 1500 python tools/run_configs.py 4000 2>&1 | grep -v "^cfg3 subset" | python -c "
import sys,json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:200]); continue
    print(d['config'], 'k',d['k'],'K',d['k_eff'],'spilled',d['spilled_groups'],'device_ms %.3f'%d['device_ms'],'G obs/s %.2f'%(d['obs_per_s_device']/1e9),'e2e_ms %.1f'%d['e2e_wall_ms'],'parity',d['parity_first_groups'],'cpu_s',d['cpu_oracle_s'])
"
