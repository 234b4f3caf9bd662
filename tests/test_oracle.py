"""CPU tests (-m "not gpu"): the oracle against the golden vectors, its two implementations against each other, and
the contig-level expectations of the reference's own unit tests (reference src/fracture.rs:593-707)."""
import numpy as np
import pytest

from util import groups_to_arrow, load_golden, random_groups


def _run(oracle, reads, k, mc, **kw):
    d, off, v = oracle.strings_to_arrow(reads)
    return oracle.count_prune_batch(d, off, np.array([0, len(reads)]), k, mc, validity=v, **kw)


def test_golden_vectors(oracle):
    gold = load_golden()
    assert len(gold["cases"]) >= 10
    for c in gold["cases"]:
        b = _run(oracle, c["reads"], c["k"], c["min_coverage"], emit_csr=True)
        assert b.k_eff == c["k_eff"]
        assert b.n_observations == c["n_observations"]
        rows = oracle.batch_group_as_tuples(b, 0)
        assert [r[0] for r in rows] == c["kmers"]
        assert [r[1] for r in rows] == c["exts"]
        assert [r[2] for r in rows] == c["counts"]
        assert oracle.digest(b, 0) == c["sha256"]
        assert [int(x) for x in b.stats[0]] == c["stats"]
        assert [int(x) for x in b.csr_row_ptr] == c["csr_row_ptr"]
        assert [int(x) for x in b.csr_col] == c["csr_col"]


def test_survey_appendix_b_facts(oracle):
    """Facts SURVEY.md Appendix B derived independently from the reference's test inputs."""
    g = {(c["name"], c["k"], c["min_coverage"]): c for c in load_golden()["cases"]}
    b1 = g[("B1", 13, 1)]
    assert b1["k_eff"] == 16 and b1["n_observations"] == 36 and len(b1["kmers"]) == 29
    assert sorted(b1["counts"]).count(2) == 7 and b1["stats"] == [29, 2, 0]
    assert g[("B1", 13, 2)]["kmers"] == sorted(set(
        "GGGCTGGTGGGCGTCCGTCTGC"[i:i + 16] for i in range(7)))
    # dangling exts survive pruning unchanged (SURVEY D-1)
    i = g[("B1", 13, 2)]["kmers"].index("GGGCTGGTGGGCGTCC")
    assert g[("B1", 13, 2)]["exts"][i] == 0x48
    b2 = g[("B2", 20, 1)]
    assert b2["k_eff"] == 32 and b2["n_observations"] == 378 and len(b2["kmers"]) == 142 and max(b2["counts"]) == 4
    assert len(g[("B2", 20, 2)]["kmers"]) == 117 and len(g[("B2", 20, 3)]["kmers"]) == 73
    b3 = g[("B3", 4, 1)]
    assert dict(zip(b3["kmers"], zip(b3["exts"], b3["counts"])))["AAAA"] == (0x33, 3)
    assert len(g[("B4", 4, 1)]["kmers"]) == 37 and sum(g[("B4", 4, 1)]["counts"]) == 60


def test_reference_unit_test_expectations(oracle):
    """reference src/fracture.rs:621-679 (44-bp answer, comment :611) and :681-707 (absent anchors -> empty)."""
    exp = load_golden()["reference_test_expectations"]
    e = exp["test_full_assembly_with_path_finding"]
    b = _run(oracle, e["reads"], e["k"], e["min_coverage"])
    nodes = oracle.batch_group_as_tuples(b, 0)
    contig = oracle.shortest_path_contig(nodes, b.k_eff, e["start_anchor"], e["end_anchor"])
    assert contig == e["expected_contig"]
    assert e["start_anchor"] in contig and e["end_anchor"] in contig
    assert oracle.compress_unitigs(nodes, b.k_eff) == {e["expected_contig"]}
    e = exp["test_assembly_with_invalid_anchors"]
    b = _run(oracle, e["reads"], e["k"], e["min_coverage"])
    assert oracle.shortest_path_contig(oracle.batch_group_as_tuples(b, 0), b.k_eff, e["start_anchor"], e["end_anchor"]) is None


def test_unitig_restatements_agree(oracle):
    """compress_unitigs (walks from the smallest unused k-mer in both directions) and compress_unitigs_full (heads first, then
    cycles) give the same contig set; the full form also folds counts and exts (reference src/fracture.rs:351-383)."""
    rng = np.random.default_rng(77)
    for k, mc in ((5, 1), (9, 1), (13, 2), (17, 1), (33, 1)):
        groups = random_groups(rng, 30, max_reads=14, max_len=100, tmpl_len=90)
        groups.append(["ACGTACGTACGTACGTACGTACGT"] * 3)                  # pure cycle
        groups.append(["AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA"])     # self-loop k-mer
        data, offs, validity, goff = groups_to_arrow(groups)
        b = oracle.count_prune_batch(data, offs, goff, k, mc, validity=validity)
        for g in range(len(groups)):
            nodes = oracle.batch_group_as_tuples(b, g)
            full = oracle.compress_unitigs_full(nodes, b.k_eff)
            assert set(full) == oracle.compress_unitigs(nodes, b.k_eff), (k, mc, g)
            assert sum(len(c) - b.k_eff + 1 for c in full) == len(nodes)   # every node is in exactly one unitig
    e = load_golden()["reference_test_expectations"]["test_full_assembly_with_path_finding"]
    b = _run(oracle, e["reads"], e["k"], e["min_coverage"])
    full = oracle.compress_unitigs_full(oracle.batch_group_as_tuples(b, 0), b.k_eff)
    assert list(full) == [e["expected_contig"]] and full[e["expected_contig"]][1] == 0   # one chain, no dangling exts


@pytest.mark.parametrize("k,exact", [(3, False), (4, False), (7, False), (13, False), (17, False), (31, False), (33, False),
                                     (64, False), (5, True), (21, True), (47, True), (1, True)])
def test_two_oracles_agree(oracle, k, exact):
    rng = np.random.default_rng(1000 + k)
    for mc in (0, 1, 2, 3):
        groups = random_groups(rng, 25, max_reads=10, max_len=90, tmpl_len=120)
        data, offs, validity, goff = groups_to_arrow(groups)
        b = oracle.count_prune_batch(data, offs, goff, k, mc, validity=validity, exact_k=exact)
        for g, reads in enumerate(groups):
            assert oracle.batch_group_as_tuples(b, g) == oracle.count_prune_group_py(reads, k, mc, exact_k=exact), (k, mc, g)


def test_k_eff_and_estimate_k(oracle):
    # reference src/fracture.rs:246-256
    assert [oracle.k_eff(k) for k in (0, 1, 4, 5, 8, 9, 16, 17, 31, 32, 33, 47, 63, 64, 65, 1000)] == \
        [4, 4, 4, 8, 8, 16, 16, 32, 32, 32, 64, 64, 64, 64, -1, -1]
    # reference src/fracture.rs:24-54
    assert oracle.estimate_k(np.array([0])) == 31                       # no sequences
    assert oracle.estimate_k(np.array([0, 0, 0])) == 31                 # all empty
    assert oracle.estimate_k(np.array([0, 150, 300])) == 49             # 150/3 = 50 -> even -> 49
    assert oracle.estimate_k(np.array([0, 300])) == 63                  # 100 -> 99 -> clamp 63
    assert oracle.estimate_k(np.array([0, 20])) == 11                   # 7 -> clamp 11
    assert oracle.estimate_k(np.array([0, 1])) == 63                    # round(0.33)=0 -> 0-1 wraps -> clamp 63
    assert oracle.estimate_k(np.array([0, 93, 93])) == 31               # empty rows excluded from the mean


def test_k_above_64_is_empty(oracle):
    b = _run(oracle, ["ACGT" * 40], 65, 1)   # reference src/fracture.rs:212-215
    assert b.k_eff == -1 and len(b.kmers_lo) == 0 and list(b.kmer_offsets) == [0, 0]


def test_saturation(oracle):
    b = _run(oracle, ["A" * 70000], 16, 1)
    assert list(b.counts) == [65535] and b.n_observations == 70000 - 15 and list(b.exts) == [0x11]


def test_threads_do_not_change_results(oracle):
    rng = np.random.default_rng(7)
    groups = random_groups(rng, 200, max_reads=8, max_len=80)
    data, offs, validity, goff = groups_to_arrow(groups)
    a = oracle.count_prune_batch(data, offs, goff, 17, 2, validity=validity, n_threads=1)
    b = oracle.count_prune_batch(data, offs, goff, 17, 2, validity=validity, n_threads=4)
    from util import assert_batches_equal
    assert_batches_equal(a, b)
