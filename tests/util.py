"""Shared helpers for the parity tests: input builders and the bit-exact comparison GPU batch <-> oracle batch."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "appendix_b.json")


def load_golden():
    with open(GOLDEN) as f:
        return json.load(f)


def groups_to_arrow(groups):
    """groups: list of list[str|None] -> (data, offsets int64, validity|None, group_offsets)"""
    from oracle.dbg_oracle import strings_to_arrow
    flat = [r for g in groups for r in g]
    data, offs, validity = strings_to_arrow(flat)
    goff = np.zeros(len(groups) + 1, dtype=np.int64)
    np.cumsum([len(g) for g in groups], out=goff[1:])
    return data, offs, validity, goff


def random_groups(rng, n_groups, max_reads=12, max_len=120, tmpl_len=200, err=0.02, p_n=0.03, p_lower=0.1, p_null=0.03,
                  p_empty=0.03, alphabet="ACGT"):
    """Groups of noisy reads drawn from a per-group random template, with the edge cases of reference
    src/fracture.rs:218-229 / src/expressions.rs:737-740 sprinkled in."""
    groups = []
    for _ in range(n_groups):
        T = int(rng.integers(max(8, tmpl_len // 4), tmpl_len + 1))
        tmpl = rng.integers(0, len(alphabet), size=T)
        reads = []
        for _ in range(int(rng.integers(0, max_reads + 1))):
            u = rng.random()
            if u < p_null:
                reads.append(None)
                continue
            if u < p_null + p_empty:
                reads.append("")
                continue
            L = int(rng.integers(1, min(max_len, T) + 1))
            s = int(rng.integers(0, T - L + 1))
            b = tmpl[s:s + L].copy()
            flip = rng.random(L) < err
            b[flip] = rng.integers(0, len(alphabet), size=int(flip.sum()))
            read = "".join(alphabet[x] for x in b)
            if rng.random() < p_lower:
                read = read.lower()
            if rng.random() < p_n:
                pos = int(rng.integers(0, L))
                read = read[:pos] + rng.choice(list("NnXRY-* ")) + read[pos + 1:]
            reads.append(read)
        groups.append(reads)
    return groups


def assert_batches_equal(gpu, ref, check_stats=True, check_csr=False, ctx=""):
    assert gpu.k_eff == ref.k_eff, ctx
    assert gpu.n_observations == ref.n_observations, f"{ctx}: n_observations {gpu.n_observations} != {ref.n_observations}"
    np.testing.assert_array_equal(gpu.kmer_offsets, ref.kmer_offsets, err_msg=f"{ctx}: kmer_offsets")
    np.testing.assert_array_equal(gpu.kmers_lo, ref.kmers_lo, err_msg=f"{ctx}: kmers_lo")
    if ref.kmers_hi is not None:
        assert gpu.kmers_hi is not None, ctx
        np.testing.assert_array_equal(gpu.kmers_hi, ref.kmers_hi, err_msg=f"{ctx}: kmers_hi")
    np.testing.assert_array_equal(gpu.exts, ref.exts, err_msg=f"{ctx}: exts")
    np.testing.assert_array_equal(gpu.counts, ref.counts, err_msg=f"{ctx}: counts")
    if check_stats and gpu.stats is not None:
        np.testing.assert_array_equal(gpu.stats, ref.stats, err_msg=f"{ctx}: stats")
    if check_csr:
        np.testing.assert_array_equal(gpu.csr_row_ptr, ref.csr_row_ptr, err_msg=f"{ctx}: csr_row_ptr")
        np.testing.assert_array_equal(gpu.csr_col, ref.csr_col, err_msg=f"{ctx}: csr_col")


def check_invariants(b, min_cov):
    """Size-independent properties of a result: strictly ascending k-mers inside each group, counts >=
    max(1,min_cov), kmer_offsets monotone."""
    off = b.kmer_offsets
    assert off[0] == 0 and np.all(np.diff(off) >= 0) and off[-1] == len(b.kmers_lo)
    if len(b.kmers_lo) == 0:
        return
    assert int(b.counts.min()) >= max(1, min_cov)
    lo = b.kmers_lo
    if b.kmers_hi is None:
        asc = lo[1:] > lo[:-1]
    else:
        hi = b.kmers_hi
        asc = (hi[1:] > hi[:-1]) | ((hi[1:] == hi[:-1]) & (lo[1:] > lo[:-1]))
    boundary = np.zeros(len(lo) - 1, dtype=bool)
    inner = off[1:-1]
    inner = inner[(inner > 0) & (inner < len(lo))]
    boundary[inner - 1] = True
    assert np.all(asc | boundary), "k-mers not strictly ascending inside a group"
