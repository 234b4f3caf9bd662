"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes -> librogtk_gpu.so), against the oracle
on identical inputs.  Bit-exact: k-mers, exts bytes, u16 counts, group offsets, node statistics and successor CSR."""
import numpy as np
import pytest

from util import assert_batches_equal, check_invariants, groups_to_arrow, load_golden, random_groups

pytestmark = pytest.mark.gpu

EMIT_CSR, EMIT_STATS, EXACT_K, FORCE_SPILL, NO_MULTIPASS = 1, 2, 4, 32, 64
NO_PARTITION = 256


def _both(engine, oracle, data, offs, goff, k, mc, validity=None, exact=False, spill=False, csr=True, ctx="", extra_flags=0):
    flags = EMIT_STATS | (EMIT_CSR if csr else 0) | (EXACT_K if exact else 0) | (FORCE_SPILL if spill else 0) | extra_flags
    got = engine.count_prune(data, offs, goff, k, mc, validity=validity, flags=flags)
    ref = oracle.count_prune_batch(data, offs, goff, k, mc, validity=validity, exact_k=exact, emit_csr=csr)
    assert_batches_equal(got, ref, check_csr=csr, ctx=ctx)
    return got, ref


@pytest.mark.parametrize("spill", [False, True])
def test_golden_fixtures(engine, oracle, spill):
    for c in load_golden()["cases"]:
        data, offs, validity = oracle.strings_to_arrow(c["reads"])
        goff = np.array([0, len(c["reads"])])
        got, _ = _both(engine, oracle, data, offs, goff, c["k"], c["min_coverage"], validity=validity, spill=spill,
                       ctx=f"{c['name']} k={c['k']} mc={c['min_coverage']}")
        rows = oracle.batch_group_as_tuples(got, 0)
        assert [r[0] for r in rows] == c["kmers"] and [r[1] for r in rows] == c["exts"] and [r[2] for r in rows] == c["counts"]
        assert oracle.digest(got, 0) == c["sha256"]
        assert [int(x) for x in got.stats[0]] == c["stats"]
        assert [int(x) for x in got.csr_col] == c["csr_col"]


@pytest.mark.parametrize("spill", [False, True])
@pytest.mark.parametrize("k,exact", [(3, False), (8, False), (13, False), (17, False), (31, False), (33, False), (47, False),
                                     (64, False), (5, True), (21, True), (27, True), (40, True), (63, True), (1, True)])
def test_random_groups(engine, oracle, k, exact, spill):
    rng = np.random.default_rng(4242 + k + (7 if spill else 0))
    for mc in (0, 1, 2, 3):
        groups = random_groups(rng, 60, max_reads=14, max_len=150, tmpl_len=220)
        data, offs, validity, goff = groups_to_arrow(groups)
        _both(engine, oracle, data, offs, goff, k, mc, validity=validity, exact=exact, spill=spill, ctx=f"k={k} mc={mc}")


def test_empty_and_degenerate_batches(engine, oracle):
    e = np.zeros(0, np.uint8)
    # no groups, no rows
    got = engine.count_prune(e, np.array([0]), np.array([0]), 17, 2, flags=EMIT_STATS | EMIT_CSR)
    assert got.n_groups == 0 and len(got.kmers_lo) == 0
    # groups without rows, rows without bases, reads shorter than K, only invalid reads, only nulls
    groups = [[], [""], ["ACG"], ["ACGTNACGTACGTACGTACGTACGTACGTACGTACGTACGT"], [None, None], [], ["acgtacgtacgtacgtacgtacgtacgtacgtacgt"]]
    data, offs, validity, goff = groups_to_arrow(groups)
    for spill in (False, True):
        got, ref = _both(engine, oracle, data, offs, goff, 17, 1, validity=validity, spill=spill)
        assert list(np.diff(got.kmer_offsets)[:6]) == [0, 0, 0, 0, 0, 0] and got.kmer_offsets[-1] > 0
    # k > 64 -> empty result, not an error (reference src/fracture.rs:212-215)
    got = engine.count_prune(data, offs, goff, 65, 1, validity=validity, flags=EMIT_STATS)
    assert got.k_eff == -1 and len(got.kmers_lo) == 0 and got.n_groups == len(groups)


def test_int32_offsets_and_unaligned_slices(engine, oracle):
    rng = np.random.default_rng(5)
    groups = random_groups(rng, 40, max_reads=10, max_len=100, p_null=0.0)
    data, offs, validity, goff = groups_to_arrow(groups)
    ref = oracle.count_prune_batch(data, offs, goff, 17, 1, validity=validity)
    got = engine.count_prune(data, offs.astype(np.int32), goff, 17, 1, validity=validity, flags=EMIT_STATS)
    assert_batches_equal(got, ref)
    # an Arrow slice: the data pointer is not 16-byte aligned and offsets do not start at 0
    pad = np.concatenate([np.frombuffer(b"NNNNN", dtype=np.uint8), data])
    got = engine.count_prune(pad, offs + 5, goff, 17, 1, validity=validity, flags=EMIT_STATS)
    assert_batches_equal(got, ref)


def test_utf8view_layout(engine, oracle):
    """Polars String columns arrive as Arrow Utf8View: inline strings (<=12 bytes) and out-of-line strings
    spread over two data buffers."""
    rng = np.random.default_rng(9)
    groups = random_groups(rng, 30, max_reads=10, max_len=60, tmpl_len=80, p_null=0.05)
    flat = [r for g in groups for r in g]
    goff = np.zeros(len(groups) + 1, dtype=np.int64)
    np.cumsum([len(g) for g in groups], out=goff[1:])
    bufs = [bytearray(), bytearray()]
    views = np.zeros((len(flat), 16), dtype=np.uint8)
    valid = np.ones(len(flat), dtype=bool)
    for i, r in enumerate(flat):
        if r is None:
            valid[i] = False
            continue
        b = r.encode()
        views[i, 0:4] = np.frombuffer(np.uint32(len(b)).tobytes(), dtype=np.uint8)
        if len(b) <= 12:
            views[i, 4:4 + len(b)] = np.frombuffer(b, dtype=np.uint8)
        else:
            which = i % 2
            off = len(bufs[which])
            bufs[which] += b
            views[i, 4:8] = np.frombuffer(b[:4], dtype=np.uint8)
            views[i, 8:12] = np.frombuffer(np.uint32(which).tobytes(), dtype=np.uint8)
            views[i, 12:16] = np.frombuffer(np.uint32(off).tobytes(), dtype=np.uint8)
    validity = None if valid.all() else np.packbits(valid, bitorder="little")
    data, offs, v2, _ = groups_to_arrow(groups)
    for k, mc in ((8, 1), (17, 2)):
        ref = oracle.count_prune_batch(data, offs, goff, k, mc, validity=v2)
        got = engine.count_prune_views(views, [np.frombuffer(bytes(b), dtype=np.uint8) for b in bufs], goff, k, mc,
                                       validity=validity, flags=EMIT_STATS)
        assert_batches_equal(got, ref, ctx=f"view k={k}")


def test_views_and_pageable_input_on_the_chunked_path(oracle, monkeypatch):
    """Large host input that is not page-locked (numpy arrays, like the buffers Polars / pyarrow own) and the Utf8View layout go
    through the bounce buffers of the overlapped path (a staging thread gathers views into the offsets layout one chunk ahead):
    many small chunks, nulls, inline and out-of-line rows over three data buffers, CSR + stats, and a malformed view in a late chunk."""
    import rogtk_b200
    monkeypatch.setenv("ROGTK_CHUNK_BYTES", "16384")
    monkeypatch.setenv("ROGTK_STAGE_THREADS", "3")
    rng = np.random.default_rng(21)
    groups = random_groups(rng, 300, max_reads=14, max_len=90, tmpl_len=120, p_null=0.05)
    flat = [r for g in groups for r in g]
    goff = np.zeros(len(groups) + 1, dtype=np.int64)
    np.cumsum([len(g) for g in groups], out=goff[1:])
    bufs = [bytearray(), bytearray(), bytearray()]
    views = np.zeros((len(flat), 16), dtype=np.uint8)
    valid = np.ones(len(flat), dtype=bool)
    for i, r in enumerate(flat):
        if r is None:
            valid[i] = False
            views[i, 0:4] = np.frombuffer(np.uint32(77).tobytes(), dtype=np.uint8)   # a null row's view is never dereferenced
            views[i, 8:12] = np.frombuffer(np.uint32(9).tobytes(), dtype=np.uint8)
            continue
        b = r.encode()
        views[i, 0:4] = np.frombuffer(np.uint32(len(b)).tobytes(), dtype=np.uint8)
        if len(b) <= 12:
            views[i, 4:4 + len(b)] = np.frombuffer(b, dtype=np.uint8)
        else:
            which = (i * 7) % 3
            off = len(bufs[which])
            bufs[which] += b
            views[i, 4:8] = np.frombuffer(b[:4], dtype=np.uint8)
            views[i, 8:12] = np.frombuffer(np.uint32(which).tobytes(), dtype=np.uint8)
            views[i, 12:16] = np.frombuffer(np.uint32(off).tobytes(), dtype=np.uint8)
    validity = None if valid.all() else np.packbits(valid, bitorder="little")
    nbufs = [np.frombuffer(bytes(b), dtype=np.uint8) for b in bufs]
    data, offs, v2, _ = groups_to_arrow(groups)
    with rogtk_b200.Engine(device=0) as eng:
        for k, mc in ((8, 1), (17, 2), (40, 1)):
            ref = oracle.count_prune_batch(data, offs, goff, k, mc, validity=v2, emit_csr=True)
            got = eng.count_prune_views(views, nbufs, goff, k, mc, validity=validity, flags=EMIT_STATS | EMIT_CSR)
            assert_batches_equal(got, ref, check_csr=True, ctx=f"chunked views k={k}")
            got = eng.count_prune(data, offs, goff, k, mc, validity=v2, flags=EMIT_STATS | EMIT_CSR)      # pageable offsets layout
            assert_batches_equal(got, ref, check_csr=True, ctx=f"chunked pageable k={k}")
            got = eng.count_prune(data, offs.astype(np.int32), goff, k, mc, validity=v2, flags=EMIT_STATS)
            assert_batches_equal(got, ref, ctx=f"chunked pageable int32 offsets k={k}")
        bad = views.copy()
        j = max(i for i, r in enumerate(flat) if r is not None and len(r) > 12)
        bad[j, 12:16] = np.frombuffer(np.uint32(1 << 30).tobytes(), dtype=np.uint8)
        with pytest.raises(rogtk_b200.RogtkGpuError) as ei:
            eng.count_prune_views(bad, nbufs, goff, 17, 2, validity=validity)
        assert ei.value.code == -1
        ref = oracle.count_prune_batch(data, offs, goff, 17, 2, validity=v2)
        assert_batches_equal(eng.count_prune_views(views, nbufs, goff, 17, 2, validity=validity, flags=EMIT_STATS), ref, ctx="after error")


def test_count_saturation_and_long_reads(engine, oracle):
    # one k-mer observed > 65535 times (u16 saturation); read longer than any staging tile
    reads = ["A" * 70000, "ACGT" * 3000, "A" * 40 + "C" + "A" * 40]
    data, offs, validity = oracle.strings_to_arrow(reads)
    got, ref = _both(engine, oracle, data, offs, np.array([0, 3]), 16, 1, ctx="saturation")
    assert got.counts.max() == 65535
    got, ref = _both(engine, oracle, data, offs, np.array([0, 1, 2, 3]), 31, 2, ctx="saturation K=32")
    # the all-T k-mer is the table's EMPTY sentinel for K=32 / K=64: must still be counted
    reads = ["T" * 100, "T" * 40 + "G" + "T" * 70, "t" * 64]
    data, offs, validity = oracle.strings_to_arrow(reads)
    for k in (16, 32, 64):
        for spill in (False, True):
            _both(engine, oracle, data, offs, np.array([0, 2, 3]), k, 1, spill=spill, ctx=f"all-T k={k}")


def test_groups_too_big_for_one_table(engine, oracle):
    """skewed groups (cfg3 shape): > 256 rows, > 16384 bases, or too many distinct k-mers for one shared-memory table are
    counted in several passes over slabs of their reads, or (ROGTK_NO_MULTIPASS / very large groups) through the HBM tables;
    the answer must not depend on the path taken."""
    from rogtk_b200 import synth
    rp = np.array([1, 2, 3, 40, 300, 7, 120, 600, 55, 1, 260, 90])
    data, offs, goff = synth.synth_reads(synth.config_seed(3), 0, len(rp), rp, 150, 500)
    for k in (11, 17, 33):
        got, ref = _both(engine, oracle, data, offs, goff, k, 2, ctx=f"skew k={k}")
        assert got.n_multipass_groups >= 3 and got.n_spilled_groups == 0
        got, ref = _both(engine, oracle, data, offs, goff, k, 2, ctx=f"skew k={k} hbm", extra_flags=NO_MULTIPASS)
        assert got.n_spilled_groups >= 3 and got.n_multipass_groups == 0
    # many distinct k-mers in few rows: 60 unrelated reads of 250 bases
    rng = np.random.default_rng(11)
    reads = ["".join("ACGT"[x] for x in rng.integers(0, 4, 250)) for _ in range(60)]
    data, offs, validity = oracle.strings_to_arrow(reads)
    got, ref = _both(engine, oracle, data, offs, np.array([0, 60]), 31, 1, ctx="distinct overflow")
    assert got.n_multipass_groups == 1 and got.n_spilled_groups == 0
    got, ref = _both(engine, oracle, data, offs, np.array([0, 60]), 31, 1, ctx="distinct overflow hbm", extra_flags=NO_MULTIPASS)
    assert got.n_spilled_groups == 1


@pytest.mark.parametrize("k,exact", [(31, False), (16, False), (47, False), (33, True), (6, True), (1, True)])
def test_multipass_groups(engine, oracle, k, exact):
    """mid-size groups: 60..430 reads of 300 bp around a 1000 bp template (1..8 passes), a group of unrelated reads (every pass
    near its table limit), null / invalid / short reads inside slabs, and a group beyond the multi-pass limit (partition-then-count;
    the HBM tables with ROGTK_NO_PARTITION)."""
    from rogtk_b200 import synth
    rp = np.array([60, 5, 130, 280, 0, 430, 90, 600])
    data, offs, goff = synth.synth_reads(synth.config_seed(2), 7, len(rp), rp, 300, 1000)
    for mc in (1, 2, 3):
        got, ref = _both(engine, oracle, data, offs, goff, k, mc, exact=exact, csr=(mc == 2), ctx=f"multipass k={k} mc={mc}")
        assert got.n_multipass_groups >= 5
        assert got.n_spilled_groups == 0           # 600 x 300 bases is beyond 8 passes: partition-then-count
        alt = engine.count_prune(data, offs, goff, k, mc, flags=EMIT_STATS | NO_PARTITION | (EXACT_K if exact else 0))
        assert alt.n_spilled_groups == 1
        assert_batches_equal(alt, ref, ctx=f"multipass k={k} mc={mc}, HBM tables for the largest group")
    rng = np.random.default_rng(k)
    groups = []
    # unrelated reads: ~ as many distinct k-mers as observations
    groups.append(["".join("ACGT"[x] for x in rng.integers(0, 4, 200)) for _ in range(300)])
    # dirty slabs: nulls, empties, non-ACGT bytes, lower case, reads shorter than k, one long read
    tmpl = "".join("ACGT"[x] for x in rng.integers(0, 4, 3000))
    dirty = []
    for i in range(400):
        s = int(rng.integers(0, 2700))
        r = tmpl[s:s + int(rng.integers(1, 300))]
        u = rng.random()
        if u < 0.05:
            r = None
        elif u < 0.10:
            r = ""
        elif u < 0.15:
            r = r[:len(r) // 2] + "N" + r[len(r) // 2 + 1:]
        elif u < 0.25:
            r = r.lower()
        dirty.append(r)
    dirty.insert(200, tmpl)
    groups.append(dirty)
    # low-complexity group: every k-mer of the group falls into very few ranges
    groups.append(["A" * 250 + "C" * 50, "T" * 300, "ACAC" * 75] * 40)
    d2, o2, v2, g2 = groups_to_arrow(groups)
    for mc in (1, 2):
        got, ref = _both(engine, oracle, d2, o2, g2, k, mc, validity=v2, exact=exact, ctx=f"multipass dirty k={k} mc={mc}")
        assert got.n_multipass_groups >= 2
    if k == 31:
        # the threshold is compared with the saturated u16 count: 500 x 219 observations of the all-A 32-mer
        d3, o3, v3, g3 = groups_to_arrow([["A" * 250] * 500])
        got, _ = _both(engine, oracle, d3, o3, g3, k, 65535, ctx="saturated count meets min_coverage 65535")
        assert got.n_multipass_groups == 1 and len(got.kmers_lo) == 1 and got.counts[0] == 65535
        got, _ = _both(engine, oracle, d3, o3, g3, k, 65536, ctx="saturated count below min_coverage 65536")
        assert len(got.kmers_lo) == 0


def test_device_generator_matches_numpy(engine):
    from rogtk_b200 import synth
    seed = synth.config_seed(2)
    d_data, d_off, d_goff, nbytes = engine.synth_reads_device(seed, 3, 50, 7, 150, 500, 655)
    data, offs, goff = synth.synth_reads(seed, 3, 50, 7, 150, 500, 655)
    assert np.array_equal(engine.d2h(d_data, nbytes), data)
    assert np.array_equal(engine.d2h(d_off, (50 * 7 + 1) * 8, np.int64), offs)
    assert np.array_equal(engine.d2h(d_goff, 51 * 8, np.int64), goff)
    for p in (d_data, d_off, d_goff):
        engine.device_free(p)
    # the UMIs of a hash shard (bench.py, N > 1) and a read offset (the ranks of the global-spectrum mode): same generator
    from rogtk_b200 import sharding
    pop = np.arange(400, dtype=np.int64)
    ids = pop[sharding.shard_ids(pop, 4) == 2][:40]
    assert len(ids) == 40 and all(engine._lib.rogtk_gpu_shard_of(int(u), 4) == 2 for u in ids[:10])
    d_data, d_off, d_goff, nbytes = engine.synth_reads_device(seed, 0, 40, 7, 150, 500, 655, umi_ids=ids)
    got = engine.d2h(d_data, nbytes).reshape(40, 7 * 150)
    for i, u in enumerate(ids):
        assert np.array_equal(got[i], synth.synth_reads(seed, int(u), 1, 7, 150, 500, 655)[0]), u
    for p in (d_data, d_off, d_goff):
        engine.device_free(p)
    d_data, d_off, d_goff, nbytes = engine.synth_reads_device(seed, 9, 5, 3, 150, 500, 655, read_base=4)
    got = engine.d2h(d_data, nbytes).reshape(5, 3 * 150)
    full = synth.synth_reads(seed, 9, 5, 7, 150, 500, 655)[0].reshape(5, 7 * 150)
    assert np.array_equal(got, full[:, 4 * 150:])
    for p in (d_data, d_off, d_goff):
        engine.device_free(p)


@pytest.mark.parametrize("cfg", ["cfg1", "cfg2"])
def test_baseline_config_subsets(engine, oracle, cfg):
    """First UMIs of BASELINE.json configs[0] (10k x 20 x 150, k=17, min_cov=2) and configs[1] (1M x 50 x 300, k=31)."""
    from rogtk_b200 import synth
    if cfg == "cfg1":
        n, R, L, T, k = 1000, 20, 150, 500, 17
    else:
        n, R, L, T, k = 300, 50, 300, 1000, 31
    data, offs, goff = synth.synth_reads(synth.config_seed(1 if cfg == "cfg1" else 2), 0, n, R, L, T)
    got, ref = _both(engine, oracle, data, offs, goff, k, 2, csr=False, ctx=cfg)
    assert got.k_eff == 32 and got.n_spilled_groups == 0
    assert got.n_observations == n * R * (L - 31)
    check_invariants(got, 2)
    _both(engine, oracle, data, offs, goff, k, 1, csr=True, ctx=cfg + " mc=1")


@pytest.mark.parametrize("k", [11, 16, 17, 31, 33, 63])
def test_baseline_config3_subset(engine, oracle, k):
    """BASELINE configs[2]: skewed reads per UMI (log-uniform 1..5000; the first 40 UMIs hold groups of 3028, 2120 and 3962
    reads), L=150, T=500, min_cov=2, k sweep over K = 16 / 32 / 64 -- single table, multi-pass and partition-then-count in one batch,
    and the same batch with the largest groups in the HBM tables (ROGTK_NO_PARTITION)."""
    from rogtk_b200 import synth
    seed = synth.config_seed(3)
    rp = synth.skewed_reads_per_umi(seed, 40)
    data, offs, goff = synth.synth_reads(seed, 0, len(rp), rp, 150, 500)
    got, ref = _both(engine, oracle, data, offs, goff, k, 2, csr=(k == 31), ctx=f"cfg3 k={k}")
    assert got.n_spilled_groups == 0 and got.n_multipass_groups >= 8
    check_invariants(got, 2)
    alt = engine.count_prune(data, offs, goff, k, 2, flags=EMIT_STATS | NO_PARTITION)
    assert alt.n_spilled_groups >= 3 and alt.n_multipass_groups >= 5
    assert_batches_equal(alt, ref, ctx=f"cfg3 k={k}, HBM tables")


def test_device_resident_path_and_invariants_at_scale(engine, oracle):
    """Inputs generated in HBM, outputs left in HBM (the `value` path of bench.py) at 20k UMIs of the cfg2 shape:
    size-independent properties + a sampled comparison against the oracle."""
    from rogtk_b200 import synth
    seed = synth.config_seed(2)
    n, R, L, T = 20000, 50, 300, 1000
    d_data, d_off, d_goff, nbytes = engine.synth_reads_device(seed, 0, n, R, L, T, 655)
    db = engine.count_prune_device(d_data, nbytes, d_off, n * R, d_goff, n, 31, 2, flags=EMIT_STATS)
    assert db.n_observations == n * R * (L - 31) and db.n_spilled_groups == 0
    got = db.to_host()
    db.free()
    check_invariants(got, 2)
    assert np.array_equal(got.stats[:, 0], np.diff(got.kmer_offsets))
    # second run gives the identical answer (no dependence on scheduling order)
    db2 = engine.count_prune_device(d_data, nbytes, d_off, n * R, d_goff, n, 31, 2, flags=EMIT_STATS)
    got2 = db2.to_host()
    db2.free()
    assert_batches_equal(got2, got)
    # sampled oracle comparison: UMIs 0..199 and 19800..19999
    for base in (0, n - 200):
        data, offs, goff = synth.synth_reads(seed, base, 200, R, L, T)
        ref = oracle.count_prune_batch(data, offs, goff, 31, 2)
        a, b = int(got.kmer_offsets[base]), int(got.kmer_offsets[base + 200])
        assert np.array_equal(got.kmer_offsets[base:base + 201] - a, ref.kmer_offsets)
        assert np.array_equal(got.kmers_lo[a:b], ref.kmers_lo) and np.array_equal(got.exts[a:b], ref.exts)
        assert np.array_equal(got.counts[a:b], ref.counts)
    for p in (d_data, d_off, d_goff):
        engine.device_free(p)


def test_full_size_headline_config(engine, oracle):
    """BASELINE configs[1] at its full size (1M UMIs x 50 x 300 bp, k=31, min_cov=2), device-resident in and out as in
    bench.py: size-independent properties on the whole result, the first half of the batch re-counted on its own, and 96
    UMIs sampled across the batch compared bit for bit with the oracle (the generator is counter-based, so any UMI's reads
    can be regenerated on the host)."""
    from rogtk_b200 import synth
    seed = synth.config_seed(2)
    n, R, L, T = 1_000_000, 50, 300, 1000
    d_data, d_off, d_goff, nbytes = engine.synth_reads_device(seed, 0, n, R, L, T, 655)
    try:
        db = engine.count_prune_device(d_data, nbytes, d_off, n * R, d_goff, n, 31, 2)
        assert db.n_observations == n * R * (L - 31) and db.n_spilled_groups == 0 and db.n_multipass_groups == 0 and db.k_eff == 32
        off = engine.d2h(db._c.kmer_offsets, (n + 1) * 8, np.int64)
        per = np.diff(off)
        assert off[0] == 0 and off[-1] == db.n_kmers and per.min() > 0
        # every consensus 32-mer is covered ~13x: at least the 969 - (a few masked by errors) survive; error k-mers seen twice add a few
        assert 900 <= per.mean() <= 1100 and per.max() < 8192
        # the first half of the batch on its own: same offsets (groups are independent)
        half = n // 2
        db2 = engine.count_prune_device(d_data, half * R * L, d_off, half * R, d_goff, half, 31, 2)
        off2 = engine.d2h(db2._c.kmer_offsets, (half + 1) * 8, np.int64)
        assert np.array_equal(off2, off[:half + 1])
        # sampled groups: counts, exts, k-mers bit for bit (and identical between the two runs)
        rng = np.random.default_rng(5)
        picks = sorted(set([0, 1, half - 1, half, n - 1] + [int(x) for x in rng.integers(0, n, 91)]))
        for g in picks:
            a, b = int(off[g]), int(off[g + 1])
            lo = engine.d2h(db._c.kmers_lo + a * 8, (b - a) * 8, np.uint64)
            ex = engine.d2h(db._c.exts + a, b - a, np.uint8)
            ct = engine.d2h(db._c.counts + a * 2, (b - a) * 2, np.uint16)
            data, offs, goff = synth.synth_reads(seed, g, 1, R, L, T)
            ref = oracle.count_prune_batch(data, offs, goff, 31, 2)
            assert np.array_equal(lo, ref.kmers_lo) and np.array_equal(ex, ref.exts) and np.array_equal(ct, ref.counts), g
            assert np.all(lo[1:] > lo[:-1]) and ct.min() >= 2
            if g < half:
                assert np.array_equal(engine.d2h(db2._c.kmers_lo + a * 8, (b - a) * 8, np.uint64), lo)
        db2.free()
        db.free()
    finally:
        for p in (d_data, d_off, d_goff):
            engine.device_free(p)


def test_chunked_overlapped_host_path(oracle, monkeypatch):
    """Large host batches are cut into chunks whose H2D / compute / D2H overlap (rogtk_gpu.cu run_chunked).  Force tiny chunks
    and compare every output array, including CSR and statistics whose indices must be re-based across chunks."""
    import rogtk_b200
    from rogtk_b200 import synth
    monkeypatch.setenv("ROGTK_CHUNK_BYTES", "65536")
    eng = rogtk_b200.Engine(device=0)
    try:
        rng = np.random.default_rng(77)
        groups = random_groups(rng, 900, max_reads=12, max_len=140, tmpl_len=200)
        data, offs, validity, goff = groups_to_arrow(groups)
        assert data.size > 3 * 65536
        for k, mc, o32 in ((17, 2, False), (31, 1, True), (40, 1, False)):
            o = offs.astype(np.int32) if o32 else offs
            got = eng.count_prune(data, o, goff, k, mc, validity=validity, flags=EMIT_STATS | EMIT_CSR)
            ref = oracle.count_prune_batch(data, offs, goff, k, mc, validity=validity, emit_csr=True)
            assert_batches_equal(got, ref, check_csr=True, ctx=f"chunked k={k}")
        # BASELINE configs[1] shape, 400 UMIs = 6 MB -> ~90 chunks
        data, offs, goff = synth.synth_reads(synth.config_seed(2), 0, 400, 50, 300, 1000)
        got = eng.count_prune(data, offs, goff, 31, 2, flags=EMIT_STATS)
        ref = oracle.count_prune_batch(data, offs, goff, 31, 2)
        assert_batches_equal(got, ref, ctx="chunked cfg2")
    finally:
        eng.close()


def test_global_spectrum_single_rank_and_emulated_sources(engine, oracle):
    """Global k-mer-spectrum mode: the merge of per-source partial (k-mer, count, exts) lists must equal one
    un-grouped filter_kmers call.  Sources are emulated on one GPU: every "rank" counts its stripe of the reads, the lists are
    split by owner and concatenated per owner exactly as the exchange would deliver them."""
    from rogtk_b200 import spectrum, synth
    data, offs, _ = synth.synth_reads(synth.config_seed(5), 0, 40, 25, 150, 500)   # 1000 reads from 40 templates
    n_rows = len(offs) - 1

    def to_dev(a):
        a = np.ascontiguousarray(a)
        p = engine.device_alloc(max(a.nbytes, 16))
        if a.nbytes:
            engine.h2d(p, a.view(np.uint8))
        return p

    for k, mc in ((31, 2), (13, 1), (40, 2)):
        ref = oracle.count_prune_batch(data, offs, np.array([0, n_rows]), k, mc)
        # single rank (rogtk_gpu_spectrum_exchange without a communicator)
        st = {}
        got = spectrum.global_spectrum(engine, data, offs, k, mc, flags=EMIT_STATS, stats_out=st)
        assert_batches_equal(got, ref, ctx=f"spectrum 1 rank k={k}")
        assert st["bytes_sent"] == 0 and st["rounds"] == 1
        # the same in rounds of 130 reads: the share of the spectrum is carried in HBM between the rounds
        got = spectrum.global_spectrum(engine, data, offs, k, mc, flags=EMIT_STATS, chunk_rows=130)
        assert_batches_equal(got, ref, ctx=f"spectrum 1 rank, 8 rounds, k={k}")
        # 3 emulated ranks, reads striped round-robin
        world = 3
        parts = []
        for r in range(world):
            rows = np.arange(r, n_rows, world)
            d = np.concatenate([data[offs[i]:offs[i + 1]] for i in rows])
            o = np.concatenate([[0], np.cumsum(offs[rows + 1] - offs[rows])]).astype(np.int64)
            b = engine.count_prune(d, o, np.array([0, len(rows)]), k, 1)
            owner = np.array([engine._lib.rogtk_gpu_spectrum_owner(int(b.kmers_lo[i]), int(b.kmers_hi[i]) if b.kmers_hi is not None else 0,
                                                                   b.k_eff, world) for i in range(len(b.kmers_lo))])
            assert np.all(np.diff(owner) >= 0)            # owners are monotone in the ascending list
            parts.append((b, owner))
        pieces = []
        for dst in range(world):
            sel = [(b, owner == dst) for b, owner in parts]
            lo = np.concatenate([b.kmers_lo[m] for b, m in sel])
            hi = np.concatenate([b.kmers_hi[m] for b, m in sel]) if ref.kmers_hi is not None else None
            ct = np.concatenate([b.counts[m] for b, m in sel])
            ex = np.concatenate([b.exts[m] for b, m in sel])
            ptrs = [to_dev(lo), to_dev(hi) if hi is not None else 0, to_dev(ct), to_dev(ex)]
            pieces.append(engine.spectrum_merge(ptrs[0] if len(lo) else 0, ptrs[1] if hi is not None and len(lo) else 0,
                                                ptrs[2] if len(lo) else 0, ptrs[3] if len(lo) else 0, len(lo), ref.k_eff, mc))
            for p in ptrs:
                if p:
                    engine.device_free(p)
        assert np.array_equal(np.concatenate([p.kmers_lo for p in pieces]), ref.kmers_lo)
        if ref.kmers_hi is not None:
            assert np.array_equal(np.concatenate([p.kmers_hi for p in pieces]), ref.kmers_hi)
        assert np.array_equal(np.concatenate([p.exts for p in pieces]), ref.exts)
        assert np.array_equal(np.concatenate([p.counts for p in pieces]), ref.counts)


def test_global_spectrum_merge_tables_and_their_fallback(engine, oracle):
    """The merge of the exchange (ascending runs -> one shared-memory table per k-mer range, kernels_part.cuh k_part_count<W, true>)
    on a spectrum large enough for hundreds of ranges, in several rounds (carry + own slice = two runs), and on reads whose k-mers
    crowd one range (12 000 distinct k-mers share their first 16 bases): the range does not fit a table, the sort-based merge
    takes over, the result is the same."""
    from rogtk_b200 import spectrum, synth
    data, offs, _ = synth.synth_reads(synth.config_seed(5), 3, 1500, 8, 150, 500)    # 12 000 reads, 1500 templates: ~1.4 M observations
    n_rows = len(offs) - 1
    for k, mc in ((31, 2), (45, 1)):
        ref = oracle.count_prune_batch(data, offs, np.array([0, n_rows]), k, mc)
        for chunk in (None, 5000):
            got = spectrum.global_spectrum(engine, data, offs, k, mc, flags=EMIT_STATS, chunk_rows=chunk)
            assert_batches_equal(got, ref, ctx=f"spectrum merge k={k} chunk={chunk}")
    rng = np.random.default_rng(5)
    reads = ["A" * 16 + "".join(rng.choice(list("ACGT"), 22)) for _ in range(12000)]
    d2, o2, _ = oracle.strings_to_arrow(reads)
    ref = oracle.count_prune_batch(d2, o2, np.array([0, len(reads)]), 31, 1)
    for chunk in (None, 7000):
        got = spectrum.global_spectrum(engine, d2, o2, 31, 1, flags=EMIT_STATS, chunk_rows=chunk)
        assert_batches_equal(got, ref, ctx=f"spectrum merge, crowded range, chunk={chunk}")


def test_threshold_reuse_matches_recount(engine, oracle):
    """one count at min_coverage=1 serves every higher threshold of the sweep / optimise drivers
    (reference src/expressions.rs:901-941, src/fracture_opt.rs:163-218) bit-for-bit."""
    rng = np.random.default_rng(123)
    groups = random_groups(rng, 80, max_reads=30, max_len=150, tmpl_len=200, err=0.03)
    data, offs, validity, goff = groups_to_arrow(groups)
    for k in (12, 31, 50):
        base = engine.count_prune_to_device(data, offs, goff, k, 1, validity=validity)
        try:
            for mc in (1, 2, 3, 5, 9, 70000):
                got = engine.rethreshold(base, mc, flags=EMIT_STATS | EMIT_CSR)
                ref = oracle.count_prune_batch(data, offs, goff, k, mc, validity=validity, emit_csr=True)
                assert_batches_equal(got, ref, check_csr=True, ctx=f"rethreshold k={k} mc={mc}")
        finally:
            base.free()


def test_low_complexity_groups_crowded_sort_buckets(engine, oracle):
    """Many surviving k-mers sharing their leading bases (homopolymer-rich reads) overflow the counting-sort buckets of
    k_sort_small and must take the bitonic kernel; order and content must not change."""
    rng = np.random.default_rng(2024)
    groups = []
    for g in range(12):
        reads = []
        for _ in range(40):
            tail = "".join("ACGT"[x] for x in rng.integers(0, 4, 40))
            reads.append("A" * int(rng.integers(30, 60)) + tail + "A" * int(rng.integers(5, 40)) + tail[:20])
        groups.append(reads)
    data, offs, validity, goff = groups_to_arrow(groups)
    for k, mc in ((31, 1), (16, 1), (40, 1), (31, 2)):
        _both(engine, oracle, data, offs, goff, k, mc, validity=validity, ctx=f"low complexity k={k} mc={mc}")


def test_malformed_input_is_rejected_not_faulted(engine):
    """Bad offsets / group offsets come back as ROGTK_ERR_INVALID_ARG; the context stays usable."""
    import rogtk_b200
    data = np.frombuffer(b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT", dtype=np.uint8)
    good_off = np.array([0, 20, 40], dtype=np.int64)
    cases = [
        (np.array([0, 30, 20], dtype=np.int64), np.array([0, 2])),      # offsets decrease
        (np.array([0, 20, 4000], dtype=np.int64), np.array([0, 2])),    # offsets beyond the data buffer
        (good_off, np.array([0, 5])),                                    # group beyond the rows
        (good_off, np.array([0, 2, 1])),                                 # group offsets decrease
        (np.array([-5, 20, 40], dtype=np.int64), np.array([0, 2])),     # negative offset
    ]
    for offs, goff in cases:
        with pytest.raises(rogtk_b200.RogtkGpuError) as ei:
            engine.count_prune(data, offs, goff, 17, 1)
        assert ei.value.code == -1
    got = engine.count_prune(data, good_off, np.array([0, 2]), 8, 1)
    assert got.n_observations == 26


def test_malformed_input_on_the_chunked_path(oracle, monkeypatch):
    """The overlapped host path cuts the batch with the caller's offsets before any kernel sees them: bad offsets at a group
    boundary are rejected on the host, bad offsets inside a chunk on the device, and the context stays usable (its page-locked
    blocks go back to the cache)."""
    import rogtk_b200
    from rogtk_b200 import synth
    monkeypatch.setenv("ROGTK_CHUNK_BYTES", "8192")
    data, offs, goff = synth.synth_reads(synth.config_seed(2), 3, 40, 20, 150, 500)
    with rogtk_b200.Engine(device=0) as eng:
        ref = oracle.count_prune_batch(data, offs, goff, 31, 2)
        assert_batches_equal(eng.count_prune(data, offs, goff, 31, 2, flags=EMIT_STATS), ref, ctx="chunked, clean")
        bad = []
        o = offs.copy(); o[goff[20]] = o[goff[20]] - 7000; bad.append(o)          # decreases at a group boundary
        o = offs.copy(); o[goff[30]:] += 1 << 40; bad.append(o)                    # beyond the data buffer from a boundary on
        o = offs.copy(); o[goff[10] + 3] = o[goff[10] + 2] - 1; bad.append(o)      # decreases inside a chunk
        o = offs.copy(); o[-1] = len(data) + 5; bad.append(o)                      # last offset beyond the buffer
        for o in bad:
            with pytest.raises(rogtk_b200.RogtkGpuError) as ei:
                eng.count_prune(data, o, goff, 31, 2)
            assert ei.value.code == -1
        assert_batches_equal(eng.count_prune(data, offs, goff, 31, 2, flags=EMIT_STATS), ref, ctx="chunked, after errors")


def test_malformed_views_are_rejected(engine, oracle):
    """Utf8View rows that name a data buffer the column does not have, or bytes beyond a buffer's end, come back as
    ROGTK_ERR_INVALID_ARG (no device fault); a null variadic buffer keeps its index."""
    import rogtk_b200
    reads = [b"ACGTACGTACGTACGTACGTAC", b"ACGT", b"TTGACCAGTACCAGTTTGACCA", b"GGGGGGGGGGGGGTTTTTTTTTTTAAAC"]
    buf0 = np.frombuffer(reads[0] + reads[2], dtype=np.uint8)
    buf2 = np.frombuffer(reads[3], dtype=np.uint8)

    def view(r, b=0, off=0):
        v = np.zeros(16, np.uint8)
        v[:4] = np.frombuffer(np.uint32(len(r)).tobytes(), np.uint8)
        if len(r) <= 12:
            v[4:4 + len(r)] = np.frombuffer(r, np.uint8)
        else:
            v[4:8] = np.frombuffer(r[:4], np.uint8)
            v[8:12] = np.frombuffer(np.uint32(b).tobytes(), np.uint8)
            v[12:16] = np.frombuffer(np.uint32(off).tobytes(), np.uint8)
        return v

    good = np.stack([view(reads[0], 0, 0), view(reads[1]), view(reads[2], 0, len(reads[0])), view(reads[3], 2, 0)])
    goff = np.array([0, 2, 4])
    data = np.frombuffer(b"".join(reads), dtype=np.uint8)
    offs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    ref = oracle.count_prune_batch(data, offs, goff, 8, 1)
    # buffer 1 is an empty (null) variadic buffer that keeps its place
    got = engine.count_prune_views(good, [buf0, np.zeros(0, np.uint8), buf2], goff, 8, 1, flags=EMIT_STATS)
    assert_batches_equal(got, ref, ctx="views with an empty buffer in the middle")
    for bad_view in (view(reads[3], 3, 0), view(reads[3], 2, 5), view(reads[3], 1, 0), view(reads[0], 0, 1 << 31)):
        bad = good.copy()
        bad[3] = bad_view
        with pytest.raises(rogtk_b200.RogtkGpuError) as ei:
            engine.count_prune_views(bad, [buf0, np.zeros(0, np.uint8), buf2], goff, 8, 1)
        assert ei.value.code == -1
    got = engine.count_prune_views(good, [buf0, np.zeros(0, np.uint8), buf2], goff, 8, 1, flags=EMIT_STATS)
    assert_batches_equal(got, ref, ctx="views after errors")


def test_arrow_list_of_strings_entry(engine, oracle):
    """a List(String) column (one row per UMI) goes through in one call; list offsets are the group offsets.
    Covers String, LargeString and StringView children, null strings, null lists and a sliced array."""
    import pyarrow as pa
    rng = np.random.default_rng(31)
    groups = random_groups(rng, 50, max_reads=9, max_len=90, tmpl_len=120, p_null=0.06)
    data, offs, validity, goff = groups_to_arrow(groups)
    ref = oracle.count_prune_batch(data, offs, goff, 17, 1, validity=validity)
    for child_t, list_ctor in ((pa.string(), pa.list_), (pa.large_string(), pa.large_list), (pa.string_view(), pa.list_)):
        arr = pa.array(groups, type=list_ctor(child_t))
        got = engine.count_prune_arrow(arr, 17, 1, flags=EMIT_STATS)
        assert_batches_equal(got, ref, ctx=f"arrow {child_t}")
    # a sliced list array and a null list
    arr = pa.array(groups, type=pa.list_(pa.string())).slice(7, 30)
    d2, o2, v2, g2 = groups_to_arrow(groups[7:37])
    assert_batches_equal(engine.count_prune_arrow(arr, 31, 2, flags=EMIT_STATS), oracle.count_prune_batch(d2, o2, g2, 31, 2, validity=v2),
                         ctx="arrow slice")
    with_null = list(groups[:10]) + [None] + list(groups[10:20])
    arr = pa.array(with_null, type=pa.list_(pa.string()))
    d3, o3, v3, g3 = groups_to_arrow(groups[:10] + [[]] + groups[10:20])
    assert_batches_equal(engine.count_prune_arrow(arr, 17, 1, flags=EMIT_STATS), oracle.count_prune_batch(d3, o3, g3, 17, 1, validity=v3),
                         ctx="arrow null list")


def test_concurrent_contexts_from_host_threads(oracle):
    """Polars calls the plugin concurrently from its rayon workers (SURVEY 8b "Threading"): one context per host thread must
    give independent, correct results; a shared context must serialise safely."""
    import threading
    import rogtk_b200
    rng = np.random.default_rng(99)
    jobs = []
    for t in range(4):
        groups = random_groups(rng, 120, max_reads=12, max_len=120, tmpl_len=160)
        data, offs, validity, goff = groups_to_arrow(groups)
        jobs.append((data, offs, validity, goff, oracle.count_prune_batch(data, offs, goff, 31, 2, validity=validity)))
    results, errors = [None] * 4, []
    shared = rogtk_b200.Engine(device=0)

    def work(i, own_ctx):
        try:
            eng = rogtk_b200.Engine(device=0) if own_ctx else shared
            d, o, v, g, _ = jobs[i]
            for _ in range(5):
                results[i] = eng.count_prune(d, o, g, 31, 2, validity=v, flags=EMIT_STATS)
            if own_ctx:
                eng.close()
        except Exception as e:   # pragma: no cover
            errors.append(e)

    for own_ctx in (True, False):
        th = [threading.Thread(target=work, args=(i, own_ctx)) for i in range(4)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errors, errors
        for i in range(4):
            assert_batches_equal(results[i], jobs[i][4], ctx=f"thread {i} own_ctx={own_ctx}")
    shared.close()


def test_many_tiny_groups(engine, oracle):
    """100 k groups of 1-3 short reads (the low end of the skewed config): per-group bookkeeping, empty groups included."""
    rng = np.random.default_rng(5150)
    n = 100_000
    reads_per = rng.integers(0, 4, size=n)
    lens = rng.integers(20, 60, size=int(reads_per.sum()))
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    data = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(offs[-1]))]
    goff = np.concatenate([[0], np.cumsum(reads_per)]).astype(np.int64)
    for k, mc in ((17, 1), (31, 1), (12, 2)):
        _both(engine, oracle, data, offs, goff, k, mc, csr=False, ctx=f"tiny groups k={k}")


@pytest.mark.parametrize("k,exact", [(31, False), (15, False), (64, False), (27, True), (45, True)])
def test_three_counting_paths_agree_on_skewed_batches(engine, k, exact):
    """One shared-memory table, several shared-memory passes, HBM tables: bit-identical batches on randomized skewed inputs
    (group sizes 0..900 reads, read lengths 0..400, per-group error rates 0..10 %, a few unrelated-read groups), larger than
    what the oracle-checked tests cover."""
    rng = np.random.default_rng(1234 + k)
    chunks, lens, rows_per_group = [], [], []
    for g in range(160):
        u = rng.random()
        n_reads = int(rng.integers(0, 12)) if u < 0.3 else int(np.exp(rng.uniform(0, np.log(900))))
        T = int(rng.integers(40, 1500))
        tmpl = rng.integers(0, 4, size=T, dtype=np.uint8)
        err = rng.uniform(0, 0.1)
        unrelated = rng.random() < 0.05
        for _ in range(n_reads):
            L = int(rng.integers(0, min(400, T) + 1))
            if unrelated:
                b = rng.integers(0, 4, size=L, dtype=np.uint8)
            else:
                s0 = int(rng.integers(0, T - L + 1))
                b = tmpl[s0:s0 + L].copy()
                flip = rng.random(L) < err
                b[flip] = rng.integers(0, 4, size=int(flip.sum()), dtype=np.uint8)
            a = np.frombuffer(b"ACGT", dtype=np.uint8)[b]
            if L and rng.random() < 0.02:
                a = a.copy()
                a[int(rng.integers(0, L))] = ord("N")
            chunks.append(a)
            lens.append(L)
        rows_per_group.append(n_reads)
    data = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    goff = np.concatenate([[0], np.cumsum(rows_per_group)]).astype(np.int64)
    base = EMIT_STATS | EMIT_CSR | (EXACT_K if exact else 0)
    for mc in (1, 2, 5):
        a = engine.count_prune(data, offs, goff, k, mc, flags=base)
        b = engine.count_prune(data, offs, goff, k, mc, flags=base | NO_MULTIPASS)
        c = engine.count_prune(data, offs, goff, k, mc, flags=base | FORCE_SPILL)
        assert a.n_multipass_groups > 0 and b.n_multipass_groups == 0 and b.n_spilled_groups >= a.n_multipass_groups
        assert c.n_spilled_groups == 160
        for other, name in ((b, "no multipass"), (c, "hbm tables only")):
            assert_batches_equal(other, a, check_csr=True, ctx=f"paths k={k} mc={mc}: {name}")
        check_invariants(a, mc)


def test_partition_then_count_tier(engine, oracle):
    """Groups beyond the multi-pass kernel (more than 131 072 bases: UMIs with thousands of reads, BASELINE configs[2]) are cut
    into k-mer ranges, their observations scattered to range buckets in HBM and every bucket counted in shared memory
    (kernels_part.cuh).  Bit-exact against the oracle for K = 16 / 32 / 64 and small K (one range), nulls and invalid reads
    inside; identical to the HBM-table path (ROGTK_NO_PARTITION); buckets are sized exactly, so a group whose reads are one k-mer
    repeated 119 000 times (count saturates at 65535) stays in this tier."""
    from rogtk_b200 import synth
    rp = np.array([1200, 3, 2500, 950, 1, 4000, 60])
    data, offs, goff = synth.synth_reads(synth.config_seed(3), 17, len(rp), rp, 150, 500)
    data = data.copy()
    data[offs[5] + 7] = ord("N")                # an invalid read inside a partitioned group
    data[offs[1300]:offs[1301]] |= 0x20         # a lower-case read
    valid = np.ones(len(offs) - 1, dtype=bool)
    valid[[10, 2000, 4400]] = False
    validity = np.packbits(valid, bitorder="little")
    for k, mc in ((31, 2), (13, 1), (47, 2), (5, 3), (63, 1)):
        ref = oracle.count_prune_batch(data, offs, goff, k, mc, validity=validity, emit_csr=True)
        got = engine.count_prune(data, offs, goff, k, mc, validity=validity, flags=EMIT_STATS | EMIT_CSR)
        assert_batches_equal(got, ref, check_csr=True, ctx=f"partition tier k={k} mc={mc}")
        assert got.n_observations == ref.n_observations
        assert got.n_spilled_groups == 0 and got.n_multipass_groups >= 4
        alt = engine.count_prune(data, offs, goff, k, mc, validity=validity, flags=EMIT_STATS | EMIT_CSR | NO_PARTITION)
        assert alt.n_spilled_groups >= 4
        assert_batches_equal(alt, got, check_csr=True, ctx=f"partition tier vs HBM tables k={k}")
    # a low-complexity giant group: 1000 reads of poly-A (one 32-mer observed ~119 000 times -> count saturates) + a normal big group
    reads = ["A" * 150] * 1000 + ["ACGT" * 37 + "AC"] * 900
    d2, o2, _ = oracle.strings_to_arrow(reads)
    g2 = np.array([0, 1000, 1900])
    for k in (31, 40):
        ref = oracle.count_prune_batch(d2, o2, g2, k, 2)
        got = engine.count_prune(d2, o2, g2, k, 2, flags=EMIT_STATS)
        assert_batches_equal(got, ref, ctx=f"partition tier, low-complexity giant groups k={k}")
        assert got.n_spilled_groups == 0
        assert int(got.counts.max()) == 65535


def test_partition_tier_one_large_group(engine, oracle):
    """One group far beyond 128 x 8192 observations (the un-grouped column of the global-spectrum mode): thousands of k-mer ranges,
    scatter items of several slabs (kernels_part.cuh PT_PBIG / PT_SUB_BIG).  Reads drawn from many templates, so most k-mers are
    seen a few times only.  Bit-exact against the oracle, no HBM tables, next to ordinary groups in the same batch."""
    from rogtk_b200 import synth
    data, offs, _ = synth.synth_reads(synth.config_seed(5), 0, 700, 20, 150, 500)    # 14 000 reads of 700 templates
    n = len(offs) - 1
    goff = np.array([0, 30, n - 40, n - 40, n])    # a small group, the large one (1.66 M observations at K = 32), an empty one, a small one
    data = data.copy()
    data[offs[5000] + 3] = ord("N")
    for k, mc in ((31, 2), (31, 1), (47, 2), (15, 1)):
        ref = oracle.count_prune_batch(data, offs, goff, k, mc, emit_csr=(k == 31))
        got = engine.count_prune(data, offs, goff, k, mc, flags=EMIT_STATS | (EMIT_CSR if k == 31 else 0))
        assert_batches_equal(got, ref, check_csr=(k == 31), ctx=f"one large group k={k} mc={mc}")
        assert got.n_observations == ref.n_observations
        assert got.n_spilled_groups == 0, (k, mc)


def _check_unitigs(engine, oracle, data, offs, goff, k, mc, min_length, validity=None, ctx=""):
    ref = oracle.count_prune_batch(data, offs, goff, k, mc, validity=validity)
    got = engine.assemble(data, offs, goff, k, mc, min_length=min_length, validity=validity)
    # the two implementations (one CTA per group in shared memory / node-parallel kernels) must give identical batches
    alt = engine.assemble(data, offs, goff, k, mc, min_length=min_length, validity=validity, node_parallel=True)
    for f in ("unitig_offsets", "seq_offsets", "seq", "counts", "exts", "cyclic", "largest", "largest_flags"):
        assert np.array_equal(getattr(got, f), getattr(alt, f)), f"{ctx}: fused vs node-parallel differ in {f}"
    assert got.n_groups == len(goff) - 1
    for g in range(got.n_groups):
        nodes = oracle.batch_group_as_tuples(ref, g)
        want = oracle.compress_unitigs_full(nodes, ref.k_eff)
        have = got.group(g)
        assert {s: (c, e, cy) for s, c, e, cy in have} == want, f"{ctx} group {g}"
        assert len(have) == len(want)
        heads = [s[:ref.k_eff] for s, _, _, _ in have]
        assert heads == sorted(heads), f"{ctx} group {g}: unitigs not listed by ascending first k-mer"
        # the plugin's answer: longest contig of at least min_length bases (reference src/fracture.rs:258-277)
        cand = [s for s in want if len(s) >= min_length]
        seq, fl = got.assembled(g)
        if not cand:
            assert seq == "" and fl == 0
        else:
            top = max(len(s) for s in cand)
            assert len(seq) == top and seq in want
            assert (fl & 1) == (1 if sum(1 for s in cand if len(s) == top) > 1 else 0)
            assert bool(fl & 2) == want[seq][2]
    return got


def test_unitig_compaction_and_plugin_contig(engine, oracle):
    """unitigs (set, counts, exts) against the oracle's restatement of compress_graph, and the contig
    assemble_sequences_expr(method="compression") returns.  The reference's own expectation for its 2-read fixture
    (44-bp contig, comment at src/fracture.rs:611) is reproduced on the device."""
    e = load_golden()["reference_test_expectations"]["test_full_assembly_with_path_finding"]
    data, offs, validity = oracle.strings_to_arrow(e["reads"])
    got = _check_unitigs(engine, oracle, data, offs, np.array([0, len(e["reads"])]), e["k"], e["min_coverage"], 0, ctx="B1")
    assert got.assembled(0) == (e["expected_contig"], 0)
    rng = np.random.default_rng(99)
    for k, mc, ml in ((5, 1, 0), (9, 1, 12), (13, 2, 0), (17, 1, 40), (31, 2, 0), (33, 1, 0), (64, 1, 70)):
        groups = random_groups(rng, 40, max_reads=14, max_len=140, tmpl_len=160)
        groups.append(["ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT"] * 3)   # pure cycles
        groups.append(["A" * 90])                                                                          # self-loop k-mer
        groups.append([])
        data, offs, validity, goff = groups_to_arrow(groups)
        _check_unitigs(engine, oracle, data, offs, goff, k, mc, ml, validity=validity, ctx=f"unitigs k={k} mc={mc}")
    # the headline shape: the consensus path, cut where an error k-mer seen twice branches off (1..19 unitigs per UMI)
    from rogtk_b200 import synth
    data, offs, goff = synth.synth_reads(synth.config_seed(2), 0, 40, 50, 300, 1000)
    got = _check_unitigs(engine, oracle, data, offs, goff, 31, 2, 200, ctx="cfg2 unitigs")
    assert all(len(got.assembled(g)[0]) >= 200 for g in range(40)) and len(got.assembled(1)[0]) == 969
    # min_coverage 1 keeps ~4500 nodes per UMI: the large instantiation of the per-group kernel; a 300-read group of unrelated
    # reads (> 6144 nodes) sends the whole batch to the node-parallel kernels
    data1, offs1, goff1 = synth.synth_reads(synth.config_seed(2), 100, 6, 50, 300, 1000)
    _check_unitigs(engine, oracle, data1, offs1, goff1, 31, 1, 100, ctx="cfg2 unitigs mc=1")
    rngu = np.random.default_rng(3)
    big = [["".join("ACGT"[x] for x in rngu.integers(0, 4, 120)) for _ in range(300)], ["ACGTTGCAAGGCTTAACCGGTTAACCGGATATCGCGATTA" * 3] * 4]
    d4, o4, v4, g4 = groups_to_arrow(big)
    _check_unitigs(engine, oracle, d4, o4, g4, 31, 1, 0, validity=v4, ctx="one group beyond a CTA")
    # k > 64: no contigs
    got = engine.assemble(data, offs, goff, 65, 2)
    assert got.k_eff == -1 and list(got.largest) == [-1] * 40


def test_assemble_arrow_batched_expression(engine, oracle):
    """The batched form of assemble_sequences_expr(method="compression"): List(String) in, one contig string per group out
    (reference src/expressions.rs:695-762 returns one String per group; "" when there is no contig)."""
    import pyarrow as pa
    rng = np.random.default_rng(21)
    groups = random_groups(rng, 30, max_reads=16, max_len=150, tmpl_len=180, p_null=0.05)
    groups[3] = None            # null list: empty group
    groups[7] = []
    for typ in (pa.list_(pa.string()), pa.large_list(pa.large_string()), pa.list_(pa.string_view())):
        arr = pa.array(groups, type=typ)
        contigs, flags = engine.assemble_arrow(arr, 17, 1, min_length=40)
        assert len(contigs) == 30 and contigs.type == pa.large_string()
        for g, reads in enumerate(groups):
            nodes = oracle.count_prune_group_py([r for r in (reads or []) if r is not None], 17, 1)
            want = [s for s in oracle.compress_unitigs(nodes, 32) if len(s) >= 40]
            got = contigs[g].as_py()
            if not want:
                assert got == "" and flags[g] == 0
            else:
                top = max(len(s) for s in want)
                assert len(got) == top and got in want
                assert (flags[g] & 1) == (1 if sum(1 for s in want if len(s) == top) > 1 else 0)


def test_pinned_cache_cap(oracle, monkeypatch):
    """ROGTK_PINNED_CACHE_BYTES bounds the page-locked blocks a context keeps between calls (0: every block is released)."""
    import rogtk_b200
    monkeypatch.setenv("ROGTK_PINNED_CACHE_BYTES", "0")
    rng = np.random.default_rng(8)
    with rogtk_b200.Engine(device=0) as eng:
        for n in (5, 40, 12):
            data, offs, validity, goff = groups_to_arrow(random_groups(rng, n))
            got = eng.count_prune(data, offs, goff, 21, 1, validity=validity, flags=EMIT_STATS)
            assert_batches_equal(got, oracle.count_prune_batch(data, offs, goff, 21, 1, validity=validity), ctx=f"cache cap n={n}")


def test_small_workspace_many_spill_waves(oracle):
    """A context created with a 1 MB workspace: the HBM-table path has to process its groups in many waves (and one group
    larger than a wave on its own); results must not change."""
    import rogtk_b200
    from rogtk_b200 import synth
    rp = np.array([30, 700, 45, 12, 260, 90, 1500, 3, 64, 410])
    data, offs, goff = synth.synth_reads(synth.config_seed(3), 11, len(rp), rp, 150, 500)
    with rogtk_b200.Engine(device=0, workspace_bytes=1 << 20) as eng:
        for k, flags in ((31, FORCE_SPILL), (47, FORCE_SPILL), (31, 0), (16, NO_MULTIPASS)):
            got = eng.count_prune(data, offs, goff, k, 2, flags=flags | EMIT_STATS | EMIT_CSR)
            ref = oracle.count_prune_batch(data, offs, goff, k, 2, emit_csr=True)
            assert_batches_equal(got, ref, check_csr=True, ctx=f"small workspace k={k} flags={flags}")


def test_survivor_pool_exhaustion_falls_back(engine, oracle, monkeypatch):
    """ROGTK_DEBUG_POOL_CAP (read when a context is created) shrinks the survivor pool of the shared-memory paths: groups that
    find no room are handed to the HBM-table path (which sizes its own storage); the batch is unchanged."""
    import rogtk_b200
    from rogtk_b200 import synth
    data, offs, goff = synth.synth_reads(synth.config_seed(2), 500, 400, 50, 300, 1000)
    ref = oracle.count_prune_batch(data, offs, goff, 31, 2)
    for cap in (1024, 20000, 150000):
        monkeypatch.setenv("ROGTK_DEBUG_POOL_CAP", str(cap))
        with rogtk_b200.Engine(device=0) as eng:
            got = eng.count_prune(data, offs, goff, 31, 2, flags=EMIT_STATS)
            assert_batches_equal(got, ref, ctx=f"pool cap {cap}")
            assert got.n_spilled_groups > 0
    monkeypatch.delenv("ROGTK_DEBUG_POOL_CAP")
    got = engine.count_prune(data, offs, goff, 31, 2, flags=EMIT_STATS)
    assert got.n_spilled_groups == 0


def test_device_cache_reuse_and_trim(oracle):
    """The context's caching allocator (host_util.cuh DevCache): batches of changing shape and flags reuse cached blocks, trim()
    returns them to the driver, results stay the oracle's before and after; two contexts on one device do not share blocks."""
    import rogtk_b200
    import torch
    from rogtk_b200 import synth
    with rogtk_b200.Engine(device=0) as a, rogtk_b200.Engine(device=0) as b:
        free0 = torch.cuda.mem_get_info(0)[0]
        for i, (n, k, mc, fl) in enumerate(((300, 31, 2, 0), (500, 17, 1, EMIT_CSR), (120, 47, 2, EMIT_STATS | EMIT_CSR), (300, 31, 2, EMIT_CSR))):
            data, offs, goff = synth.synth_reads(synth.config_seed(2), 10 * i, n, 20, 150, 400)
            ref = oracle.count_prune_batch(data, offs, goff, k, mc, emit_csr=bool(fl & EMIT_CSR))
            for eng in (a, b):
                assert_batches_equal(eng.count_prune(data, offs, goff, k, mc, flags=fl | EMIT_STATS), ref, check_csr=bool(fl & EMIT_CSR), ctx=f"round {i}")
            if i == 1:
                a.trim()
        # nothing stays allocated behind a finished call (buffers that outlive an entry point's scope included: the exchange's
        # merge input), and the steady state allocates nothing new
        from rogtk_b200 import spectrum
        data, offs, goff = synth.synth_reads(synth.config_seed(5), 0, 30, 20, 150, 400)
        for it in range(3):
            spectrum.global_spectrum(a, data, offs, 31, 2, chunk_rows=250)
            st = a.cache_stats()
            assert st["live_bytes"] == 0, st
            if it == 1:
                mallocs = st["n_malloc"]
        assert st["n_malloc"] == mallocs and st["n_reuse"] > 0
        assert b.cache_stats()["live_bytes"] == 0
        held = free0 - torch.cuda.mem_get_info(0)[0]
        a.trim()
        b.trim()
        assert free0 - torch.cuda.mem_get_info(0)[0] < max(held, 1)   # cached blocks went back


def test_device_cache_gives_way_under_memory_pressure(engine, oracle, monkeypatch):
    """Two contexts on one device (one per worker thread in the plugin): when a cudaMalloc fails, the idle blocks of EVERY context's
    cache go back to the driver and the request is retried; ROGTK_DEVICE_CACHE_BYTES bounds what a context keeps."""
    import rogtk_b200
    import torch
    from rogtk_b200 import synth
    data, offs, goff = synth.synth_reads(synth.config_seed(2), 0, 4000, 50, 300, 1000)      # 60 MB of reads, ~0.5 GB of working set
    small = synth.synth_reads(synth.config_seed(2), 0, 50, 50, 300, 1000)
    ref = oracle.count_prune_batch(*small, 31, 2)
    with rogtk_b200.Engine(device=0) as a, rogtk_b200.Engine(device=0) as b:
        a.count_prune(data, offs, goff, 31, 2)                      # a's cache now holds a few hundred MB of idle blocks
        assert a.cache_stats()["idle_bytes"] > 100 << 20
        # (the session's context is a third cache on the device: blocks go largest first, whoever holds them)
        held = a.cache_stats()["idle_bytes"] + engine.cache_stats()["idle_bytes"]
        rel0 = a.cache_stats()["n_released_for_others"] + engine.cache_stats()["n_released_for_others"]
        free_now = torch.cuda.mem_get_info(0)[0]
        hog = torch.empty(free_now - (64 << 20), dtype=torch.uint8, device="cuda:0")      # leave 64 MB: less than b needs
        try:
            got = b.count_prune(data, offs, goff, 31, 2, flags=EMIT_STATS)                  # must take a's idle memory
            assert a.cache_stats()["n_released_for_others"] + engine.cache_stats()["n_released_for_others"] > rel0
            assert a.cache_stats()["idle_bytes"] + engine.cache_stats()["idle_bytes"] < held
            assert got.n_groups == 4000
            assert_batches_equal(b.count_prune(*small, 31, 2, flags=EMIT_STATS), ref, ctx="after giving way")
        finally:
            del hog
            torch.cuda.empty_cache()
    monkeypatch.setenv("ROGTK_DEVICE_CACHE_BYTES", str(32 << 20))
    with rogtk_b200.Engine(device=0) as c:
        c.count_prune(data, offs, goff, 31, 2)
        assert c.cache_stats()["idle_bytes"] <= 32 << 20
        assert_batches_equal(c.count_prune(*small, 31, 2, flags=EMIT_STATS), ref, ctx="bounded cache")


@pytest.mark.parametrize("rl", [3, 5])
def test_link_walk_kernel_matches_oracle(oracle, monkeypatch, rl):
    """The experimental first-tier kernel k_walk_count (ROGTK_WALK_RL, read when a context is created): 8-byte
    reference slots + successor links.  Same batches as the default kernel: golden shape, skewed groups, K = 4 .. 64, nulls."""
    import rogtk_b200
    from rogtk_b200 import synth
    monkeypatch.setenv("ROGTK_WALK_RL", str(rl))
    with rogtk_b200.Engine(device=0) as eng:
        data, offs, goff = synth.synth_reads(synth.config_seed(2), 77, 300, 50, 300, 1000)
        for k, mc in ((31, 2), (47, 2), (13, 1), (4, 1), (63, 3), (8, 0)):
            got = eng.count_prune(data, offs, goff, k, mc, flags=EMIT_STATS | EMIT_CSR)
            ref = oracle.count_prune_batch(data, offs, goff, k, mc, emit_csr=True)
            assert_batches_equal(got, ref, check_csr=True, ctx=f"walk rl={rl} k={k} min_cov={mc}")
        rp = np.array([1, 2, 700, 45, 12, 260, 90, 3, 64, 5, 110])
        data, offs, goff = synth.synth_reads(synth.config_seed(3), 5, len(rp), rp, 150, 500)
        for k in (17, 33):
            got = eng.count_prune(data, offs, goff, k, 2, flags=EMIT_STATS)
            ref = oracle.count_prune_batch(data, offs, goff, k, 2)
            assert_batches_equal(got, ref, ctx=f"walk rl={rl} skewed k={k}")


def test_two_word_kmers_at_scale(engine, oracle):
    """K = 64 (k = 47 and k = 63; 128-bit keys, one CTA per SM) on 5000 UMIs of the headline shape, device-resident: invariants
    on the whole batch and the first / last 100 UMIs against the oracle."""
    from rogtk_b200 import synth
    seed = synth.config_seed(2)
    n, R, L, T = 5000, 50, 300, 1000
    d_data, d_off, d_goff, nbytes = engine.synth_reads_device(seed, 0, n, R, L, T, 655)
    try:
        for k in (47, 63):
            db = engine.count_prune_device(d_data, nbytes, d_off, n * R, d_goff, n, k, 2, flags=EMIT_STATS)
            assert db.k_eff == 64 and db.n_observations == n * R * (L - 63) and db.n_spilled_groups == 0
            got = db.to_host()
            db.free()
            check_invariants(got, 2)
            for base in (0, n - 100):
                data, offs, goff = synth.synth_reads(seed, base, 100, R, L, T)
                ref = oracle.count_prune_batch(data, offs, goff, k, 2)
                a, b = int(got.kmer_offsets[base]), int(got.kmer_offsets[base + 100])
                assert np.array_equal(got.kmer_offsets[base:base + 101] - a, ref.kmer_offsets)
                assert np.array_equal(got.kmers_lo[a:b], ref.kmers_lo) and np.array_equal(got.kmers_hi[a:b], ref.kmers_hi)
                assert np.array_equal(got.exts[a:b], ref.exts) and np.array_equal(got.counts[a:b], ref.counts)
    finally:
        for p in (d_data, d_off, d_goff):
            engine.device_free(p)


def test_many_tiny_groups_with_sentinel_kmer_through_hbm_tables(engine, oracle):
    """The all-T k-mer lives outside the HBM tables and is emitted with its group's first slots; with hundreds of tiny tables
    the survivor records must still come out grouped by group (the per-group segmented sort depends on it)."""
    rng = np.random.default_rng(17)
    groups = []
    for g in range(300):
        reads = []
        if g % 3 != 1:
            reads.append("T" * int(rng.integers(33, 70)))
        if g % 2:
            reads.append("".join("ACGT"[x] for x in rng.integers(0, 4, int(rng.integers(20, 60)))))
        if g % 5 == 0:
            reads.append("T" * 40 + "A" + "T" * 35)
        groups.append(reads)
    data, offs, validity, goff = groups_to_arrow(groups)
    for k in (16, 32, 64):
        _both(engine, oracle, data, offs, goff, k, 1, validity=validity, spill=True, ctx=f"tiny sentinel groups k={k}")
