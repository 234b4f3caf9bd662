"""diagnostics: assemble_sequences(compression) from pinned host buffers, call by call, with the device-cache counters (run on the GPU box).
usage: asm_probe.py [n_umis]"""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rogtk_b200
from rogtk_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
eng = rogtk_b200.Engine()
R, L = 50, 300
d_data, d_off, d_goff, nbytes = eng.synth_reads_device(synth.config_seed(2), 0, n, R, L, 1000, 655)
h = [eng.pinned_alloc(nbytes + 16), eng.pinned_alloc((n * R + 1) * 8), eng.pinned_alloc((n + 1) * 8)]
for dst, src, nb in ((h[0], d_data, nbytes), (h[1], d_off, (n * R + 1) * 8), (h[2], d_goff, (n + 1) * 8)):
    eng._check(eng._lib.rogtk_gpu_memcpy_d2h(eng._ctx, ctypes.c_void_p(dst), ctypes.c_void_p(src), nb))
def view(ptr, cnt, dt):
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(np.ctypeslib.as_ctypes_type(np.dtype(dt)))), shape=(cnt,))
data, offs, goff = view(h[0], nbytes, np.uint8), view(h[1], n * R + 1, np.int64), view(h[2], n + 1, np.int64)
for it in range(5):
    t0 = time.perf_counter()
    r = eng.assemble_sequences(data, offs, goff, k=31, min_coverage=2, method="compression", min_length=200)
    dt = (time.perf_counter() - t0) * 1e3
    print("assemble wall", round(dt, 1), "ms_device", round(r.ms_device, 1), eng.cache_stats(), flush=True)
for it in range(3):
    t0 = time.perf_counter()
    g = eng.count_prune(data, offs, goff, 31, 2)
    dt = (time.perf_counter() - t0) * 1e3
    print("count_prune wall", round(dt, 1), eng.cache_stats(), flush=True)
