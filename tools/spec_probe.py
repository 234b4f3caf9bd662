"""diagnostics: single-rank global_spectrum iterations with the device-cache counters (run on the GPU box).  usage: spec_probe.py n_templates [chunk_rows]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rogtk_b200
from rogtk_b200 import synth, spectrum
eng = rogtk_b200.Engine()
n_t = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
chunk = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else None
data, offs, _ = synth.synth_reads(synth.config_seed(5), 0, n_t, 50, 150, 500)
for it in range(6):
    st = {}
    t0 = time.perf_counter()
    got = spectrum.global_spectrum(eng, data, offs, 31, 2, chunk_rows=chunk, stats_out=st)
    dt = (time.perf_counter() - t0) * 1e3
    print(n_t, "wall", round(dt, 1), {k: round(v, 2) for k, v in st.items() if k.startswith("ms_")}, eng.cache_stats(), flush=True)
