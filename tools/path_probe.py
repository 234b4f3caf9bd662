"""diagnostics: the shortest-path methods on the device at scale (run on the GPU box): reads resident in HBM -> contigs on the host,
rogtk_gpu_assemble_batch through the C ABI.  usage: path_probe.py [n_umis ...]"""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rogtk_b200
from rogtk_b200 import synth, _ffi
eng = rogtk_b200.Engine()
lib, ctx = eng._lib, eng._ctx
R, L = 50, 300
for n in [int(x) for x in sys.argv[1:]] or [200000]:
    d_data, d_off, d_goff, nbytes = eng.synth_reads_device(synth.config_seed(2), 0, n, R, L, 1000, 655)
    dptr = (ctypes.c_void_p * 1)(d_data); dlen = (ctypes.c_int64 * 1)(nbytes)
    col = _ffi.StrCol(layout=_ffi.UTF8_I64, n_buffers=1, n_rows=n * R, validity=None, validity_bit_offset=0, offsets_or_views=d_off, data=dptr, data_len=dlen)
    for method in ("shortest_path_auto", "compression"):
        ap = _ffi.AssemblyParamsC(k=31, min_coverage=2, method=_ffi.METHODS[method], min_length=0, only_largest=1)
        for it in range(2):
            so = _ffi.StringBatchC()
            t0 = time.perf_counter()
            eng._check(lib.rogtk_gpu_assemble_batch(ctx, ctypes.byref(col), ctypes.c_void_p(d_goff), n, ctypes.byref(ap), _ffi.INPUT_ON_DEVICE, ctypes.byref(so)))
            dt = (time.perf_counter() - t0) * 1e3
            nb = int(so.n_bytes); msd = float(so.ms_device)
            lib.rogtk_gpu_free_strings(ctx, ctypes.byref(so))
        print(f"{n} UMIs {method}: wall {dt:.1f} ms, device {msd:.1f} ms, contig bytes {nb}, {n / dt * 1e3:.0f} groups/s", flush=True)
    for p in (d_data, d_off, d_goff):
        eng.device_free(p)
