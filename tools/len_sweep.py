"""diagnostics: cost of partially filled 128-window work units -- k_smem_count rate vs read length (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rogtk_b200
from rogtk_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
eng = rogtk_b200.Engine()
eng.phase_cycles(True)
for L in (159, 172, 287, 300, 415):
    d_data, d_off, d_goff, nbytes = eng.synth_reads_device(synth.config_seed(2), 0, n, 50, L, 1000, 655)
    best = None
    for it in range(3):
        db = eng.count_prune_device(d_data, nbytes, d_off, n * 50, d_goff, n, 31, 2)
        r = (db.ms_hot_kernel, db.n_observations, db.n_kmers); db.free()
        pc = eng.phase_cycles(True).astype(np.float64)
        if best is None or r[0] < best[0][0]:
            best = (r, pc)
    (ms, nobs, nk), pc = best
    print(f"L={L}: windows/read={L-31} units/read={-(-(L-31)//128)} kernel {ms:.2f} ms  {nobs/ms/1e6:.1f} Gobs/s  "
          f"count cycles/obs={pc[3]/nobs:.3f} total cycles/obs={pc.sum()/nobs:.3f} survivors/UMI={nk/n:.0f}", flush=True)
    for p in (d_data, d_off, d_goff):
        eng.device_free(p)
