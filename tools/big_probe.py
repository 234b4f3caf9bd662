"""diagnostics: one large group through the partition tier (run on the GPU box).  usage: big_probe.py [n_templates ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rogtk_b200
from rogtk_b200 import synth
eng = rogtk_b200.Engine()
for n_t in [int(x) for x in sys.argv[1:]] or [20000]:
    data, offs, _ = synth.synth_reads(synth.config_seed(5), 0, n_t, 50, 150, 500)
    n = len(offs) - 1
    for it in range(5):
        t0 = time.perf_counter()
        db = eng.count_prune_to_device(data, offs, np.array([0, n], dtype=np.int64), 31, 1)
        dt = (time.perf_counter() - t0) * 1e3
        print(n_t, "obs", db.n_observations, "kmers", db.n_kmers, "ms_total", round(db.ms_total, 2), {k: round(v, 2) for k, v in db.stage_ms.items()},
              "wall", round(dt, 1), "spilled", db.n_spilled_groups, eng.cache_stats(), flush=True)
        db.free()
