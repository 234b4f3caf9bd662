"""diagnostics: the headline shape (UMIs x 50 reads x 300 bp, min_cov 2) at K_eff = 16 / 32 / 64, device-resident (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rogtk_b200
from rogtk_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
eng = rogtk_b200.Engine()
d_data, d_off, d_goff, nbytes = eng.synth_reads_device(synth.config_seed(2), 0, n, 50, 300, 1000, 655)
for k in (15, 31, 47):
    best = None
    for it in range(4):
        db = eng.count_prune_device(d_data, nbytes, d_off, n * 50, d_goff, n, k, 2)
        r = (db.ms_total, db.ms_hot_kernel, db.n_observations, db.n_kmers, db.k_eff, db.n_multipass_groups, db.n_spilled_groups); db.free()
        if it and (best is None or r[0] < best[0]):
            best = r
    ms, msk, nobs, nk, ke, nm, ns = best
    print(f"k={k} K_eff={ke}: step {ms:.2f} ms (hot kernel {msk:.2f}), {nobs / ms / 1e6:.1f} G obs/s, {n / ms / 1e3:.2f} M UMIs/s, "
          f"{nk / n:.0f} survivors/UMI, multi-pass {nm}, HBM {ns}", flush=True)
