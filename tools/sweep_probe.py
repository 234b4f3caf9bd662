"""diagnostics: the sweep driver call by call with the device-cache counters (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rogtk_b200
from rogtk_b200 import synth
eng = rogtk_b200.Engine()
n = 20000
data, offs, goff = synth.synth_reads(synth.config_seed(2), 0, n, 50, 300, 1000)
for it in range(4):
    c0 = eng.cache_stats()
    t0 = time.perf_counter()
    sw = eng.sweep_assembly_params(data, offs, goff, k_start=25, k_end=40, k_step=5, cov_start=1, cov_end=6, cov_step=1, method="compression")
    dt = (time.perf_counter() - t0) * 1e3
    c1 = eng.cache_stats()
    print(f"sweep {dt:.1f} ms  device {sw['ms_device']:.1f} host {sw['ms_host']:.1f}  mallocs +{c1['n_malloc'] - c0['n_malloc']} reuse +{c1['n_reuse'] - c0['n_reuse']} idle {c1['idle_bytes'] >> 20} MB", flush=True)
