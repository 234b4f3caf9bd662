"""check: 60 000 small groups through the HBM tables (one segmented sort over 60 000 segments) == shared-memory path (run on the GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rogtk_b200
from rogtk_b200 import synth, _ffi
eng = rogtk_b200.Engine()
data, offs, goff = synth.synth_reads(synth.config_seed(1), 0, 60000, 6, 120, 300)
for k in (17, 47):
    a = eng.count_prune(data, offs, goff, k, 1, flags=_ffi.EMIT_STATS)
    b = eng.count_prune(data, offs, goff, k, 1, flags=_ffi.EMIT_STATS | _ffi.FORCE_SPILL)
    ok = all(np.array_equal(getattr(a, f), getattr(b, f)) for f in ("kmer_offsets", "kmers_lo", "exts", "counts", "stats"))
    if a.kmers_hi is not None:
        ok = ok and np.array_equal(a.kmers_hi, b.kmers_hi)
    print(f"k={k}: groups 60000 survivors {len(a.kmers_lo)} spilled {b.n_spilled_groups} equal {ok}  hbm path {b.ms_total:.2f} ms, smem path {a.ms_total:.2f} ms")
    assert ok
