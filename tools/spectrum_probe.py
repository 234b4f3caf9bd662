"""diagnostics: stage times of the global k-mer-spectrum mode on one GPU (world = 1): local count (one group spanning all
reads -> HBM-table path), owner split, merge.  Run on the GPU box."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rogtk_b200
from rogtk_b200 import synth, spectrum, _ffi
eng = rogtk_b200.Engine()
n_tmpl = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
data, offs, _ = synth.synth_reads(synth.config_seed(5), 0, n_tmpl, 50, 150, 500)
n_rows = len(offs) - 1
def t(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); return r, (time.perf_counter() - t0) * 1e3
for it in range(2):
    local, ms_count = t(lambda: eng.count_prune_to_device(data, offs, np.array([0, n_rows], dtype=np.int64), 31, 1))
    c, n = local._c, local.n_kmers
    _, ms_split = t(lambda: eng.spectrum_split(c.kmers_lo, c.kmers_hi, n, 32, 2))
    merged, ms_merge = t(lambda: eng.spectrum_merge(c.kmers_lo, 0, c.counts, c.exts, n, 32, 2))
    print(f"reads {n_rows} obs {local.n_observations} distinct {n}: count {ms_count:.1f} ms (device pipeline {local.ms_total:.1f} ms, "
          f"{local.n_observations / local.ms_total / 1e6:.2f} Gobs/s), split {ms_split:.2f} ms, merge {ms_merge:.1f} ms -> {len(merged.kmers_lo)} survivors", flush=True)
    local.free()
_, ms_all = t(lambda: spectrum.global_spectrum(eng, data, offs, 31, 2))
_, ms_chunk = t(lambda: spectrum.global_spectrum(eng, data, offs, 31, 2, chunk_rows=n_rows // 5 + 1))
print(f"global_spectrum: one round {ms_all:.1f} ms, five rounds {ms_chunk:.1f} ms")
