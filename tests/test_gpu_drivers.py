"""GPU parity of the whole-expression entries (rogtk_gpu_assemble_batch / _sweep_batch / _optimize_batch, through the C ABI) against
the oracle's literal restatements of the reference's drivers (one full recount per probe): assemble_sequences_expr and
assemble_sequences_with_anchors_expr (reference src/expressions.rs:695-849), fracture_sequences (src/fracture.rs:506-564),
sweep_assembly_params_expr (src/expressions.rs:880-955), optimize_assembly_expr (src/fracture_opt.rs:120-367)."""
import numpy as np
import pytest

from util import groups_to_arrow, random_groups

EMIT_CSR, EMIT_STATS = 1, 2

pytestmark = pytest.mark.gpu


def _template_groups(rng, n_groups, n_reads, rlen, tlen, err=0.01, alphabet="ACGT"):
    groups, templates = [], []
    for _ in range(n_groups):
        t = "".join(rng.choice(list(alphabet), tlen))
        templates.append(t)
        reads = []
        for _ in range(n_reads):
            s = int(rng.integers(0, tlen - rlen + 1))
            r = list(t[s:s + rlen])
            for i in range(len(r)):
                if rng.random() < err:
                    r[i] = str(rng.choice(list("ACGT")))
            reads.append("".join(r))
        # the two ends are always covered, so that anchored paths exist
        reads.append(t[:rlen])
        reads.append(t[-rlen:])
        groups.append(reads)
    return groups, templates


def test_reference_fixture_through_the_expression_entry(engine):
    """reference src/fracture.rs:611-679 (44-bp answer) and :681-707 (absent anchors: empty string, not an error)."""
    import rogtk_b200
    reads = ["GAGACTGCATGGGCTGGTGGGCGTCCGTCTGC", "GGGCTGGTGGGCGTCCGTCTGCTTTAGTGAGGGT"]
    data, offs, val, goff = groups_to_arrow([reads, reads[:1], []])
    want = "GAGACTGCATGGGCTGGTGGGCGTCCGTCTGCTTTAGTGAGGGT"
    r = engine.assemble_sequences(data, offs, goff, k=13, min_coverage=1, method="shortest_path", start_anchor="GAGACTGCAT",
                                  end_anchor="TTTAGTGAGGGT")
    assert r.strings() == [want, "", ""]
    r = engine.assemble_sequences(data, offs, goff, k=13, min_coverage=1, method="compression")
    assert r.strings() == [want, reads[0], ""] and r.n_count_passes == 1
    r = engine.assemble_sequences(data, offs, goff, k=13, min_coverage=1, method="shortest_path_auto")
    assert r.strings() == [want, reads[0], ""]
    r = engine.assemble_sequences(data, offs, goff, k=13, min_coverage=1, method="shortest_path", start_anchor="AAAAAAAAAAAA",
                                  end_anchor="TTTAGTGAGGGT")
    assert r.strings() == ["", "", ""]
    assert engine.assemble_sequences(data, offs, goff, k=80, min_coverage=1, method="compression").strings() == ["", "", ""]
    # the reference's argument errors (src/expressions.rs:700-731)
    for kw, msg in ((dict(method="compression", start_anchor="A"), "should not be provided for compression"),
                    (dict(method="shortest_path", start_anchor="A"), "Both start_anchor and end_anchor are required"),
                    (dict(method="shortest_path_auto", end_anchor="A"), "should not be provided for shortest_path_auto"),
                    (dict(method="shortest_path", anchors_per_group=(["A"] * 3, ["C"] * 3), start_anchor=None), None)):
        if msg is None:
            engine.assemble_sequences(data, offs, goff, k=13, min_coverage=1, **kw)   # per-group anchors + shortest_path is valid
            continue
        with pytest.raises(rogtk_b200.RogtkGpuError) as ei:
            engine.assemble_sequences(data, offs, goff, k=13, min_coverage=1, **kw)
        assert ei.value.code == -1 and msg in str(ei.value)
    with pytest.raises(rogtk_b200.RogtkGpuError):
        engine.assemble_sequences(data, offs, goff, k=13, min_coverage=1, method="compression", anchors_per_group=(["A"] * 3, ["C"] * 3))


@pytest.mark.parametrize("k,mc", [(13, 1), (17, 2), (31, 2), (40, 1), (6, 2)])
def test_assemble_sequences_all_methods(engine, oracle, k, mc):
    """Every method on noisy groups, scalar and per-group anchors, min_length, nulls / lower case / invalid reads."""
    rng = np.random.default_rng(100 + k)
    groups, templates = _template_groups(rng, 40, 14, 70, 130, err=0.015)
    groups[3] = [None, "acgtacgtnacgt", ""] + groups[3]
    groups[5] = []
    groups[7] = [r.lower() for r in groups[7]]
    data, offs, val, goff = groups_to_arrow(groups)
    kw = dict(validity=val)
    # compression (the expression's only_largest=True) + min_length
    for ml in (None, 100):
        got = engine.assemble_sequences(data, offs, goff, k=k, min_coverage=mc, method="compression", min_length=ml, **kw)
        for g, reads in enumerate(groups):
            want, info = oracle.assemble_sequences_py(reads, k, mc, "compression", min_length=ml, only_largest=True)
            if info["tie"]:
                assert got.flags[g] & 1
                continue
            assert got.strings()[g] == (want[0] if want else ""), (g, ml)
            assert bool(got.flags[g] & 2) == info["cyclic"]
    # fracture_sequences' listing: every contig, joined with newlines (set comparison: the listing order is the reference's boomphf order)
    got = engine.assemble_sequences(data, offs, goff, k=k, min_coverage=mc, method="compression", only_largest=False, min_length=40, **kw)
    for g, reads in enumerate(groups):
        want, _ = oracle.assemble_sequences_py(reads, k, mc, "compression", min_length=40, only_largest=False)
        s = got.strings()[g]
        assert sorted(s.split("\n") if s else []) == sorted(want), g
        assert (s.split("\n") if s else []) == want   # and this repo's documented order: ascending first k-mer
    # shortest_path with one anchor pair for all groups (rarely present) and with per-group anchors (the dynamic-anchor expression)
    sl = [t[:6] for t in templates]
    el = [t[-6:] for t in templates]
    got = engine.assemble_sequences(data, offs, goff, k=k, min_coverage=mc, method="shortest_path", anchors_per_group=(sl, el), **kw)
    n_found = 0
    for g, reads in enumerate(groups):
        want, _ = oracle.assemble_sequences_py(reads, k, mc, "shortest_path", sl[g], el[g], only_largest=True)
        assert got.strings()[g] == (want[0] if want else ""), g
        n_found += bool(want)
    assert n_found >= 10 or mc > 1 or k > 32   # (min_coverage 2 prunes the read ends where the anchors sit; K = 64 leaves 7 windows per read)
    got = engine.assemble_sequences(data, offs, goff, k=k, min_coverage=mc, method="shortest_path", start_anchor=sl[0], end_anchor=el[0],
                                    min_length=50, **kw)
    for g, reads in enumerate(groups):
        want, _ = oracle.assemble_sequences_py(reads, k, mc, "shortest_path", sl[0], el[0], min_length=50, only_largest=True)
        assert got.strings()[g] == (want[0] if want else ""), g
    got = engine.assemble_sequences(data, offs, goff, k=k, min_coverage=mc, method="shortest_path_auto", **kw)
    for g, reads in enumerate(groups):
        want, _ = oracle.assemble_sequences_py(reads, k, mc, "shortest_path_auto", only_largest=True)
        assert got.strings()[g] == (want[0] if want else ""), g


def test_auto_k_per_group(engine, oracle):
    """auto_k (reference src/fracture.rs:203-209): k = estimate_k(group) per group; groups of different read lengths fall into
    different K_eff classes and are counted in one pass per class."""
    rng = np.random.default_rng(7)
    groups = []
    for rlen, tlen in ((36, 80), (60, 110), (120, 200), (150, 260)):   # k = 11 (K 16), 19 (K 32), 39 (K 64), 49 (K 64)
        gs, _ = _template_groups(rng, 6, 10, rlen, tlen, err=0.005)
        groups += gs
    groups.append([])            # estimate_k of nothing = 31
    groups.append(["", None])
    data, offs, val, goff = groups_to_arrow(groups)
    got = engine.assemble_sequences(data, offs, goff, k=999, min_coverage=2, method="compression", auto_k=True, validity=val)
    assert got.n_count_passes == 3 and all(f & 4 for f in got.flags)
    for g, reads in enumerate(groups):
        want, info = oracle.assemble_sequences_py(reads, 999, 2, "compression", only_largest=True, auto_k=True)
        if info["tie"]:
            continue
        assert got.strings()[g] == (want[0] if want else ""), (g, info)


@pytest.mark.parametrize("method", ["compression", "shortest_path", "shortest_path_auto"])
def test_sweep_matches_recounting(engine, oracle, method):
    """sweep_assembly_params: the grid in the reference's order; one counting pass per distinct K_eff instead of one per cell."""
    rng = np.random.default_rng(11)
    groups, templates = _template_groups(rng, 12, 12, 60, 100, err=0.02)
    groups[2] = []
    data, offs, val, goff = groups_to_arrow(groups)
    sa, ea = (templates[0][:5], templates[0][-5:]) if method == "shortest_path" else (None, None)
    res = engine.sweep_assembly_params(data, offs, goff, k_start=3, k_end=70, k_step=9, cov_start=0, cov_end=4, cov_step=2, method=method,
                                       start_anchor=sa, end_anchor=ea, validity=val)
    ks = list(range(3, 71, 9))
    assert res["k"].shape == (len(groups), len(ks) * 3)
    assert res["n_count_passes"] == len({oracle.k_eff_py(k) for k in ks if k <= 64})
    for g, reads in enumerate(groups):
        want = oracle.sweep_py(reads, 3, 70, 9, 0, 4, 2, method, sa, ea)
        if method == "compression":
            # a tie between longest contigs has one length: comparable without the order caveat
            pass
        got = list(zip(res["k"][g].tolist(), res["min_coverage"][g].tolist(), res["contig_length"][g].tolist()))
        assert got == want, g


@pytest.mark.parametrize("method,explore_k,prioritize", [("shortest_path", False, False), ("shortest_path", True, False),
                                                         ("shortest_path", False, True), ("shortest_path", True, True)])
def test_optimize_matches_recounting(engine, oracle, method, explore_k, prioritize):
    """optimize_assembly: the hill-climb per group, every probed cell evaluated once for the whole batch."""
    rng = np.random.default_rng(13)
    groups, templates = _template_groups(rng, 10, 16, 64, 110, err=0.02)
    groups[4] = [None, "ACGTNNNN"]
    groups[6] = []
    data, offs, val, goff = groups_to_arrow(groups)
    sa, ea = templates[1][:6], templates[1][-6:]
    res = engine.optimize_assembly(data, offs, goff, sa, ea, method=method, start_k=15, start_min_coverage=4, max_iterations=6,
                                   explore_k=explore_k, prioritize_length=prioritize, validity=val)
    total_probes = 0
    for g, reads in enumerate(groups):
        (contig, k, cov, length, nseq), probes = oracle.optimize_py(reads, 15, 4, sa, ea, method, 6, explore_k, prioritize)
        total_probes += probes
        if method == "compression" and contig:
            # Compression ties are order-dependent (boomphf); lengths still agree
            assert int(res["length"][g]) == length, g
        else:
            assert res["contig"].strings()[g] == contig, g
            assert (int(res["k"][g]), int(res["min_coverage"][g]), int(res["length"][g])) == (k, cov, length), g
        assert int(res["input_sequences"][g]) == nseq
    assert res["n_probes"] >= total_probes          # (the batch evaluates the probes an early return would have skipped)
    assert res["n_cells_evaluated"] < res["n_probes"]
    # required anchors (src/fracture_opt.rs:298-302); the anchors are always handed to AssemblyMethod::from_str (:304-308), which
    # rejects them for every method but shortest_path: optimize_assembly(method="compression") is an error in the reference
    import rogtk_b200
    with pytest.raises(rogtk_b200.RogtkGpuError):
        engine.optimize_assembly(data, offs, goff, None, ea, validity=val)
    for m in ("compression", "shortest_path_auto"):
        with pytest.raises(rogtk_b200.RogtkGpuError) as ei:
            engine.optimize_assembly(data, offs, goff, sa, ea, method=m, validity=val)
        assert "should not be provided" in str(ei.value)


def test_assemble_compression_on_the_overlapped_path(oracle, monkeypatch):
    """Large host-resident input: assemble_sequences(method="compression") runs chunk by chunk (count -> unitigs -> contig choice on the
    device, only the strings come back).  Same rows as the plain route and as the oracle; only_largest False lists every contig."""
    import rogtk_b200
    from rogtk_b200 import synth
    data, offs, goff = synth.synth_reads(synth.config_seed(2), 40, 600, 20, 150, 400)
    with rogtk_b200.Engine(device=0) as eng:
        plain = {ol: eng.assemble_sequences(data, offs, goff, k=31, min_coverage=2, method="compression", min_length=60, only_largest=ol)
                 for ol in (True, False)}
    monkeypatch.setenv("ROGTK_CHUNK_BYTES", "100000")
    with rogtk_b200.Engine(device=0) as eng:
        for ol in (True, False):
            got = eng.assemble_sequences(data, offs, goff, k=31, min_coverage=2, method="compression", min_length=60, only_largest=ol)
            assert got.n_count_passes == 1
            assert got.strings() == plain[ol].strings()
            assert (got.flags == plain[ol].flags).all()
        got = eng.assemble_sequences(data, offs, goff, k=31, min_coverage=2, method="compression", min_length=60)
    s = got.strings()
    n_checked = 0
    for g in range(0, 600, 7):
        reads = [data[offs[r]:offs[r + 1]].tobytes().decode() for r in range(goff[g], goff[g + 1])]
        want, info = oracle.assemble_sequences_py(reads, 31, 2, "compression", min_length=60, only_largest=True)
        if info["tie"]:
            continue
        assert s[g] == (want[0] if want else ""), g
        n_checked += 1
    assert n_checked > 50


def test_one_frame_on_several_gpus(oracle):
    """rogtk_gpu_count_prune_multi (BASELINE configs[3] as written): whole groups to the device their UMI id hashes to, one context +
    host thread per device, no collective, results in the original group order.  Two contexts on device 0 exercise the whole path
    on a one-GPU box; with a second GPU present the same frame also runs across devices 0 and 1."""
    import rogtk_b200
    from rogtk_b200 import synth, sharding
    from util import assert_batches_equal
    rp = np.concatenate([np.full(150, 12), np.array([1, 2, 300, 45, 0, 260, 90, 3])])
    data, offs, goff = synth.synth_reads(synth.config_seed(3), 5, len(rp), rp, 150, 500)
    umi_ids = np.arange(len(rp), dtype=np.uint64) * 7919 + 13
    valid = np.ones(len(offs) - 1, dtype=bool)
    valid[::17] = False
    validity = np.packbits(valid, bitorder="little")
    devs = [[0, 0], [0, 0, 0]]
    if rogtk_b200.ffi.load().rogtk_gpu_device_count() >= 2:
        devs.append([0, 1])
    for dl in devs:
        engines = [rogtk_b200.Engine(device=d) for d in dl]
        try:
            for k, mc, fl in ((31, 2, EMIT_STATS | EMIT_CSR), (47, 1, EMIT_STATS), (80, 1, 0)):
                got, shard = rogtk_b200.count_prune_multi(engines, data, offs, goff, k, mc, umi_ids=umi_ids, validity=validity, flags=fl)
                ref = oracle.count_prune_batch(data, offs, goff, k, mc, validity=validity, emit_csr=bool(fl & EMIT_CSR))
                if k <= 64:
                    assert_batches_equal(got, ref, check_csr=bool(fl & EMIT_CSR), ctx=f"multi {dl} k={k}")
                else:
                    assert got.k_eff == -1 and len(got.kmers_lo) == 0
                assert (shard == sharding.shard_ids(umi_ids, len(dl))).all()
                assert len(set(shard.tolist())) == len(dl)
            # the group index is hashed when there are no ids
            got, shard = rogtk_b200.count_prune_multi(engines, data, offs.astype(np.int32), goff, 17, 2)
            assert_batches_equal(got, oracle.count_prune_batch(data, offs, goff, 17, 2), ctx="multi, no ids, int32 offsets")
            assert (shard == sharding.shard_ids(np.arange(len(rp), dtype=np.uint64), len(dl))).all()
            bad = offs.copy()
            bad[40] = bad[39] - 3
            with pytest.raises(rogtk_b200.RogtkGpuError) as ei:
                rogtk_b200.count_prune_multi(engines, data, bad, goff, 17, 2)
            assert ei.value.code == -1
        finally:
            for e in engines:
                e.close()


_REF_FASTA = [   # reference src/fracture.rs:573-584 (create_test_fasta)
    "ATGCATGCATGCTAGCTGATCGATCGTAGCTAGCTAGCTGATCGATCGTACGTACGTACGTAGCTACGTACGTACGTAGCTAGCTGATCGTAGCTACGTAGCTAGCTAGCTGATCGTACGTACGT",
    "GTAGCTAGCTAGCTGATCGATCGTACGTACGTACGTAGCTACGTACGTACGTAGCTAGCTGATCGTAGCTACGTAGCTAGCTAGCTGATCGTACGTACGTAGCTGATCGATCGTAGCTACGTACGT",
    "GTACGTACGTACGTAGCTACGTACGTACGTAGCTAGCTGATCGTAGCTACGTAGCTAGCTAGCTGATCGTACGTACGTAGCTGATCGATCGTAGCTACGTACGTACGTAGCTACGTACGTACGTAG",
    "TACGTACGTACGTAGCTAGCTGATCGTAGCTACGTAGCTAGCTAGCTGATCGTACGTACGTAGCTGATCGATCGTAGCTACGTACGTACGTAGCTACGTACGTACGTAGCTAGCTGATCGTAGCT",
]


def test_fracture_entries_mirror_the_pyo3_functions(engine, oracle, tmp_path):
    """fracture_sequences / fracture_fasta (reference src/fracture.rs:468-564) with the reference's own test inputs
    (src/fracture.rs:593-607 FASTA, :621-679 two reads + anchors, :681-707 absent anchors) and its error types."""
    import rogtk_b200
    reads = ["GAGACTGCATGGGCTGGTGGGCGTCCGTCTGC", "GGGCTGGTGGGCGTCCGTCTGCTTTAGTGAGGGT"]
    got = rogtk_b200.fracture_sequences(reads, 13, 1, min_length=None, method="shortest_path", start_anchor="GAGACTGCATGG",
                                        end_anchor="TTTAGTGAGGGT", export_graphs=True, auto_k=False, prefix="test_prefix", engine=engine)
    assert got == "GAGACTGCATGGGCTGGTGGGCGTCCGTCTGCTTTAGTGAGGGT"   # the 44-bp answer in the comment at src/fracture.rs:611
    assert rogtk_b200.fracture_sequences(["AAAACCCCCAAAAA", "TTTTTGGGGGTTTT"], 4, 1, min_length=None, method="shortest_path",
                                         start_anchor="NONEXISTENT", end_anchor="ALSONOTHERE", engine=engine) == ""
    # default min_length = 200 hides the 44-bp contig; compression is the default method
    assert rogtk_b200.fracture_sequences(reads, 13, 1, engine=engine) == ""
    assert rogtk_b200.fracture_sequences(reads, 13, 1, min_length=10, engine=engine) == "GAGACTGCATGGGCTGGTGGGCGTCCGTCTGCTTTAGTGAGGGT"
    with pytest.raises(ValueError, match="Both start_anchor and end_anchor are required"):
        rogtk_b200.fracture_sequences(reads, 13, 1, method="shortest_path", engine=engine)
    with pytest.raises(ValueError, match="Invalid assembly method"):
        rogtk_b200.fracture_sequences(reads, 13, 1, method="greedy", engine=engine)

    # FASTA: wrapped lines, lower case, a record with an N (dropped), a blank line
    fa = tmp_path / "reads.fa"
    with open(fa, "w") as fh:
        for i, r in enumerate(_REF_FASTA):
            fh.write(f">read{i + 1} some description\n")
            body = r.lower() if i == 1 else r
            for j in range(0, len(body), 60):
                fh.write(body[j:j + 60] + "\n")
        fh.write(">bad\nACGTNACGT\n\n")
    assert [s.upper() for s in map(bytes.decode, rogtk_b200.fracture.read_fasta_sequences(fa))][:4] == _REF_FASTA
    want_all, _ = oracle.assemble_sequences_py(_REF_FASTA + ["ACGTNACGT"], 20, 1, "compression", min_length=0, only_largest=False)
    # (the reference test asserts a first contig > 150 bp, src/fracture.rs:603-606; no reading of the crate's semantics yields
    #  one from these four reads -- tools/parity_sensitivity.py, DESIGN.md 3 -- so the unitigs of 79 and 125 bp are what is compared)
    assert sorted(len(c) for c in want_all) == [79, 125]
    want, info = oracle.assemble_sequences_py(_REF_FASTA, 20, 1, "compression", min_length=100, only_largest=True)
    assert not info["tie"]
    assert rogtk_b200.fracture_fasta(str(fa), 20, 1, min_length=100, engine=engine) == want[0]
    assert sorted(rogtk_b200.fracture_sequences(_REF_FASTA, 20, 1, min_length=0, engine=engine).split("\n")) == sorted(want_all)
    with pytest.raises(ValueError, match="No contigs found with length >= 5000"):
        rogtk_b200.fracture_fasta(str(fa), 20, 1, min_length=5000, engine=engine)
    with pytest.raises(ValueError, match="No contigs found"):
        rogtk_b200.fracture_fasta(str(fa), 70, 1, min_length=1, engine=engine)   # k > 64: no contigs (src/fracture.rs:286-289)


def test_path_search_on_the_device_equals_the_host_route(oracle, monkeypatch):
    """The shortest-path methods run per group on the device (one thread per group, pathfind_device.cuh executing pathfind_core.h);
    ROGTK_HOST_PATHFIND=1 keeps them on the host cores (the route a boomphf node_rank takes).  Same strings, for scalar and per-group
    anchors, the automatic method, K = 16 / 32 / 64, min_length, and with a workspace small enough for several waves; both equal the
    oracle's petgraph restatement."""
    import rogtk_b200
    rng = np.random.default_rng(77)
    groups, templates = _template_groups(rng, 300, 14, 90, 220, err=0.02)
    groups[17] = []
    groups[40] = [groups[40][0]]
    data, offs, val, goff = groups_to_arrow(groups)
    sl = [t[:7] for t in templates]
    el = [t[-7:] for t in templates]
    monkeypatch.setenv("ROGTK_HOST_PATHFIND", "1")
    with rogtk_b200.Engine(device=0) as host_eng:
        monkeypatch.delenv("ROGTK_HOST_PATHFIND")
        monkeypatch.setenv("ROGTK_PATHFIND_MIN_GROUPS", "0")   # (frames of fewer than 16 384 groups search on the host by default)
        with rogtk_b200.Engine(device=0) as dev_eng, rogtk_b200.Engine(device=0, workspace_bytes=1 << 20) as small_eng:
            for k, mc in ((13, 1), (21, 2), (40, 1)):
                for kw in (dict(method="shortest_path_auto"), dict(method="shortest_path", anchors_per_group=(sl, el)),
                           dict(method="shortest_path", start_anchor=templates[3][:6], end_anchor=templates[3][-6:]),
                           dict(method="shortest_path_auto", min_length=150)):
                    want = host_eng.assemble_sequences(data, offs, goff, k=k, min_coverage=mc, validity=val, **kw).strings()
                    got = dev_eng.assemble_sequences(data, offs, goff, k=k, min_coverage=mc, validity=val, **kw).strings()
                    assert got == want, (k, mc, kw.get("method"))
                    assert small_eng.assemble_sequences(data, offs, goff, k=k, min_coverage=mc, validity=val, **kw).strings() == want
                    if k == 13 and "start_anchor" not in kw:   # (at min_coverage 2 / K = 64 the template ends are rarely covered)
                        assert sum(1 for s in got if s) > 200
            # a group beyond the device kernel's node limit (4096) is searched on the host inside the same call
            big_t = "".join(rng.choice(list("ACGT"), 7000))
            big = [big_t[i:i + 150] for i in range(0, 7000 - 150, 40)] + [big_t[:150], big_t[-150:]]
            d2, o2, v2, g2 = groups_to_arrow([groups[0], big, groups[1]])
            # (a path of more than 1001 nodes does not exist for the reference: its back-trace gives up after 1000 steps, src/djfind.rs:185)
            for kw in (dict(method="shortest_path_auto"), dict(method="shortest_path", start_anchor=big_t[:8], end_anchor=big_t[-8:]),
                       dict(method="shortest_path", start_anchor=big_t[:8], end_anchor=big_t[892:900])):
                want = host_eng.assemble_sequences(d2, o2, g2, k=21, min_coverage=1, validity=v2, **kw).strings()
                got = dev_eng.assemble_sequences(d2, o2, g2, k=21, min_coverage=1, validity=v2, **kw).strings()
                assert got == want
            assert got[1] == big_t[:900]
            # and against the oracle (per-group anchors, K = 32)
            got = dev_eng.assemble_sequences(data, offs, goff, k=21, min_coverage=2, validity=val, method="shortest_path", anchors_per_group=(sl, el)).strings()
            for g, reads in enumerate(groups):
                want, _ = oracle.assemble_sequences_py([r for r in reads if r is not None], 21, 2, "shortest_path", sl[g], el[g], only_largest=True)
                assert got[g] == (want[0] if want else ""), g
