"""diagnostics: latency of ONE small group per call (the legacy one-plugin-call-per-group form), host buffers in and out (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rogtk_b200
from rogtk_b200 import synth
from oracle import dbg_oracle
eng = rogtk_b200.Engine()
for (R, L, T, k) in ((20, 150, 500, 17), (50, 300, 1000, 31)):
    data, offs, goff = synth.synth_reads(synth.config_seed(1), 0, 1, R, L, T)
    for _ in range(20):
        eng.count_prune(data, offs, goff, k, 2)
    l0 = eng.launch_count
    t0 = time.perf_counter()
    N = 500
    for _ in range(N):
        g = eng.count_prune(data, offs, goff, k, 2)
    dt = (time.perf_counter() - t0) / N
    t1 = time.perf_counter()
    for _ in range(20):
        r = dbg_oracle.count_prune_batch(data, offs, goff, k, 2, n_threads=1)
    dc = (time.perf_counter() - t1) / 20
    t2 = time.perf_counter()
    for _ in range(200):
        s = eng.assemble_sequences(data, offs, goff, k=k, min_coverage=2, method="compression")
    da = (time.perf_counter() - t2) / 200
    print(f"{R} reads x {L} bp, k={k}: count_prune {dt * 1e6:.0f} us per call ({(eng.launch_count - l0) / N / 1:.1f} launches... ), "
          f"assemble_sequences(compression) {da * 1e6:.0f} us, C oracle 1 thread {dc * 1e6:.0f} us", flush=True)
