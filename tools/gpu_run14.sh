bash tools/gpu_run4.sh 2>&1 | grep -vE "^\+" | tail -4 | cut -c1-330
timeout 1500 python tools/run_configs.py 4000 2>&1 | grep -v "^cfg3 subset" | python -c "
import sys,json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:200]); continue
    print(d['config'], 'k',d['k'],'K',d['k_eff'],'spilled',d['spilled_groups'],'device_ms %.3f'%d['device_ms'],'hot_ms %.3f'%d['hot_kernel_ms'],'G obs/s %.2f'%(d['obs_per_s_device']/1e9),'parity',d['parity_first_groups'])
"
