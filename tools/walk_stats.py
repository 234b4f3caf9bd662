"""diagnostics: event counts of k_walk_count on the cfg2 shape (run on the GPU box).  usage: walk_stats.py [n_umis] [k]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rogtk_b200
from rogtk_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 31
eng = rogtk_b200.Engine()
d_data, d_off, d_goff, nbytes = eng.synth_reads_device(synth.config_seed(2), 0, n, 50, 300, 1000, 655)
eng.phase_cycles(True)
for it in range(2):
    db = eng.count_prune_device(d_data, nbytes, d_off, n * 50, d_goff, n, k, 2)
    ms = (db.ms_total, db.ms_hot_kernel); nobs = db.n_observations; db.free()
    pc = eng.phase_cycles(True).astype(np.float64)
names = ["meta", "ingest", "prefix+clear", "count", "prune+reserve", "compact"]
print(f"RL={os.environ.get('ROGTK_WALK_RL', 'default')} k={k}: ms_kernel={ms[1]:.3f} obs/UMI={nobs / n:.0f}  cycles/UMI: " +
      "  ".join(f"{nm}={c / n:.0f}" for nm, c in zip(names, pc[:6])))
ev = ["walked", "dense", "heads", "probe_steps", "dense_warp_steps", "walk_warp_steps", "ref_compares", "broken_ranges"]
print("   per UMI: " + "  ".join(f"{nm}={c / n:.0f}" for nm, c in zip(ev, pc[8:16])))
