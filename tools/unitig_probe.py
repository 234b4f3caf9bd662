"""diagnostics: unitig compaction on the headline shape, device-resident in and out (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rogtk_b200
from rogtk_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
eng = rogtk_b200.Engine()
d_data, d_off, d_goff, nbytes = eng.synth_reads_device(synth.config_seed(2), 0, n, 50, 300, 1000, 655)
for mc in (2, 1):
    eng.count_prune_device(d_data, nbytes, d_off, n * 50, d_goff, n, 31, mc).free()   # warm-up: the memory pool grows once
    db = eng.count_prune_device(d_data, nbytes, d_off, n * 50, d_goff, n, 31, mc)
    best = None
    for it in range(3):
        nu, nb, ms = eng.compress_unitigs(db, 200, sizes_only=True)
        best = ms if best is None else min(best, ms)
    print(f"min_cov={mc}: {n} UMIs, {db.n_kmers} nodes ({db.n_kmers / n:.0f}/UMI): count+prune {db.ms_total:.2f} ms "
          f"(multi-pass groups {db.n_multipass_groups}, HBM-table groups {db.n_spilled_groups}); unitigs {best:.2f} ms "
          f"({db.n_kmers / best / 1e6:.2f} G nodes/s), {nu / n:.1f} unitigs/UMI, {nb / n:.0f} contig bases/UMI", flush=True)
    db.free()
# flags on a host-visible subset
data, offs, goff = synth.synth_reads(synth.config_seed(2), 0, 2000, 50, 300, 1000)
cb = eng.assemble(data, offs, goff, 31, 2, min_length=200)
print(f"2000 UMIs: ties in length {int((cb.largest_flags & 1).sum())}, cyclic {int((cb.largest_flags & 2).sum() // 2)}, "
      f"no contig >= 200 bp {int((cb.largest < 0).sum())}, mean plugin contig {np.mean([len(cb.assembled(g)[0]) for g in range(2000)]):.0f} bp")
