"""CPU tests: the C-ABI library loads, exports every symbol include/rogtk_gpu.h declares, the host-side helpers that
mirror reference host logic agree with the oracle, and the product path fails loudly without a GPU."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    lib = os.path.join(ROOT, "rogtk_b200", "librogtk_gpu.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
    return lib


def test_header_symbols_exported():
    _ensure_built()
    from rogtk_b200 import _ffi
    lib = _ffi.load()
    hdr = open(os.path.join(ROOT, "include", "rogtk_gpu.h")).read()
    declared = sorted(set(re.findall(r"\b(rogtk_(?:gpu|host)_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    assert sorted(_ffi.SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert lib.rogtk_gpu_abi_version() == 2


def test_host_helpers_match_oracle(oracle):
    _ensure_built()
    import rogtk_b200
    for k in list(range(0, 70)) + [100, 1000]:
        assert rogtk_b200.k_eff(k) == oracle.k_eff(k)     # reference src/fracture.rs:246-256
    rng = np.random.default_rng(3)
    for _ in range(200):
        lens = rng.integers(0, 400, size=int(rng.integers(0, 12)))
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        assert rogtk_b200.estimate_k(off) == oracle.estimate_k(off)   # reference src/fracture.rs:24-54
    # sharding is a pure function of the UMI id, stable and in range
    ids = np.arange(10000)
    for n in (1, 2, 4, 8):
        s = np.array([rogtk_b200.shard_of(int(i), n) for i in ids])
        assert s.min() >= 0 and s.max() < n
        if n > 1:
            assert np.bincount(s, minlength=n).min() > 10000 / n * 0.9


def test_no_cpu_fallback_without_gpu():
    _ensure_built()
    import rogtk_b200
    from rogtk_b200 import _ffi
    if _ffi.load().rogtk_gpu_device_count() > 0:
        pytest.skip("a GPU is visible; the loud-failure path is exercised on CPU-only boxes")
    with pytest.raises(rogtk_b200.RogtkGpuError):
        rogtk_b200.Engine()


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under rogtk_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rogtk_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "dbg_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f


def test_synth_generator_shape_and_determinism():
    from rogtk_b200 import synth
    d1, o1, g1 = synth.synth_reads(synth.config_seed(1), 5, 7, 20, 150, 500)
    d2, o2, g2 = synth.synth_reads(synth.config_seed(1), 5, 7, 20, 150, 500)
    assert np.array_equal(d1, d2) and len(d1) == 7 * 20 * 150 and o1[-1] == len(d1) and g1[-1] == 140
    assert set(np.unique(d1)) <= set(b"ACGT")
    # shifting umi_base shifts the stream: UMI 6 of base 5 == UMI 1 of base 10
    d3, _, _ = synth.synth_reads(synth.config_seed(1), 10, 2, 20, 150, 500)
    assert np.array_equal(d1[5 * 3000:6 * 3000], d3[0:3000])
    rp = synth.skewed_reads_per_umi(3, 2000)
    assert rp.min() >= 1 and rp.max() <= 5000 and np.median(rp) < 200
