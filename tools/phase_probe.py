"""diagnostics: per-phase cycle breakdown of k_smem_count on the cfg2 shape (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rogtk_b200
from rogtk_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
eng = rogtk_b200.Engine()
d_data, d_off, d_goff, nbytes = eng.synth_reads_device(synth.config_seed(2), 0, n, 50, 300, 1000, 655)
eng.phase_cycles(True)
for it in range(3):
    db = eng.count_prune_device(d_data, nbytes, d_off, n * 50, d_goff, n, 31, 2)
    ms = (db.ms_total, db.ms_hot_kernel); db.free()
    pc = eng.phase_cycles(True).astype(np.float64)
    names = ["meta", "ingest", "prefix+clear", "count", "prune+reserve", "compact", "-", "-"]
    tot = pc.sum()
    print(f"run {it}: ms_total={ms[0]:.3f} ms_kernel={ms[1]:.3f}  cycles/UMI: " +
          "  ".join(f"{nm}={c / n:.0f}" for nm, c in zip(names, pc) if nm != "-") + f"  total={tot / n:.0f}")
