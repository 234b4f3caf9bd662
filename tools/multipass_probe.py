"""diagnostics: mid-size UMI groups (more reads than one shared-memory table takes) -- several shared-memory passes vs the
HBM-table path, device-resident inputs (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rogtk_b200
from rogtk_b200 import synth, _ffi
eng = rogtk_b200.Engine()
RS = [int(x) for x in sys.argv[1:]] or [50, 60, 100, 200, 400, 600]
for R in RS:
    n = max(600, 120000 // R)
    d_data, d_off, d_goff, nbytes = eng.synth_reads_device(synth.config_seed(2), 0, n, R, 300, 1000, 655)
    row = []
    for flags in (0, _ffi.NO_MULTIPASS):
        best = None
        for it in range(3):
            db = eng.count_prune_device(d_data, nbytes, d_off, n * R, d_goff, n, 31, 2, flags=flags)
            r = (db.ms_total, db.n_observations, db.n_kmers, db.n_multipass_groups, db.n_spilled_groups); db.free()
            if best is None or r[0] < best[0]:
                best = r
        row.append(best)
    (m0, nobs, nk, nm, ns0), (m1, _, nk1, _, ns1) = row
    assert nk == nk1
    print(f"R={R:4d} n={n:5d}: multipass {m0:8.2f} ms {nobs/m0/1e6:7.1f} Gobs/s (multi {nm}, hbm {ns0}) | hbm-only {m1:8.2f} ms "
          f"{nobs/m1/1e6:7.1f} Gobs/s (hbm {ns1})  speedup {m1/m0:.2f}x  survivors/UMI {nk/n:.0f}", flush=True)
    for p in (d_data, d_off, d_goff):
        eng.device_free(p)
