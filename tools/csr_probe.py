"""diagnostics: device time of the CSR / stats stages on the cfg2 shape (run on the GPU box).  usage: csr_probe.py [n_umis] [plain_steps_first]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rogtk_b200
from rogtk_b200 import synth, _ffi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
pre = int(sys.argv[2]) if len(sys.argv) > 2 else 0
eng = rogtk_b200.Engine()
d_data, d_off, d_goff, nbytes = eng.synth_reads_device(synth.config_seed(2), 0, n, 50, 300, 1000, 655)
for it in range(pre):
    db = eng.count_prune_device(d_data, nbytes, d_off, n * 50, d_goff, n, 31, 2, flags=0)
    print("plain", round(db.ms_total, 3), flush=True)
    db.free()
for it in range(4):
    db = eng.count_prune_device(d_data, nbytes, d_off, n * 50, d_goff, n, 31, 2, flags=_ffi.EMIT_CSR | _ffi.EMIT_STATS)
    print({k: round(v, 3) for k, v in db.stage_ms.items()}, "total", round(db.ms_total, 3), "edges", db.n_edges, "nodes", db.n_kmers, flush=True)
    db.free()
