"""CPU tests of the host path-finding restatement (rogtk_b200/csrc/pathfind_host.h through rogtk_host_shortest_path, no GPU): the
C++ product code against the oracle's independent Python restatement of reference src/djfind.rs (petgraph's first-pop Dijkstra
with Rust's BinaryHeap tie order, the 1e-9 back-trace, anchor rules, the auto mode's candidates and scoring)."""
import ctypes

import numpy as np
import pytest




def _host_path(lib, b, g, start, end, rank=None):
    o, e = int(b.kmer_offsets[g]), int(b.kmer_offsets[g + 1])
    n = e - o
    lo = np.ascontiguousarray(b.kmers_lo[o:e])
    hi = np.ascontiguousarray(b.kmers_hi[o:e]) if b.kmers_hi is not None and len(b.kmers_hi) else None
    cnt = np.ascontiguousarray(b.counts[o:e])
    rp = np.ascontiguousarray(b.csr_row_ptr[o:e + 1])
    col = np.ascontiguousarray(b.csr_col)
    out = np.zeros(n + 80, np.uint8)
    rk = np.ascontiguousarray(rank, dtype=np.int32) if rank is not None else None
    r = lib.rogtk_host_shortest_path(b.k_eff, n, lo.ctypes.data, hi.ctypes.data if hi is not None else None, cnt.ctypes.data, rp.ctypes.data,
                                     col.ctypes.data, rk.ctypes.data if rk is not None else None,
                                     start.encode() if start is not None else None, end.encode() if end is not None else None,
                                     out.ctypes.data, len(out))
    assert r >= 0
    return out[:r].tobytes().decode() if r else None


@pytest.fixture(scope="module")
def lib():
    from rogtk_b200 import _ffi
    return _ffi.load()


def test_reference_fixture_path(lib, oracle):
    """reference src/fracture.rs:611-679: the 44-bp answer between the two anchors; absent anchors give nothing (:681-707)."""
    reads = ["GAGACTGCATGGGCTGGTGGGCGTCCGTCTGC", "GGGCTGGTGGGCGTCCGTCTGCTTTAGTGAGGGT"]
    data, offs, _ = oracle.strings_to_arrow(reads)
    b = oracle.count_prune_batch(data, offs, np.array([0, 2]), 13, 1, emit_csr=True)
    want = "GAGACTGCATGGGCTGGTGGGCGTCCGTCTGCTTTAGTGAGGGT"
    assert _host_path(lib, b, 0, "GAGACTGCAT", "TTTAGTGAGGGT") == want
    assert _host_path(lib, b, 0, None, None) == want
    assert _host_path(lib, b, 0, "AAAAAAAAAAAA", "TTTAGTGAGGGT") is None
    assert _host_path(lib, b, 0, "GAGACTGCAT", "CCCCCCCCCCCCCCCCCCCCCCC") is None
    # start == end node: the reference's back-trace loop never runs, path_valid stays false (src/djfind.rs:176-228)
    assert _host_path(lib, b, 0, "GAGACTGCATGGGCTG", "GAGACTGCATGGGCTG") is None


@pytest.mark.parametrize("k,alphabet,n_reads,rlen,tlen", [(13, "ACGT", 12, 60, 120), (6, "AC", 10, 40, 60), (4, "ACG", 14, 30, 40),
                                                          (33, "ACGT", 10, 90, 150), (8, "ACGT", 20, 50, 90)])
def test_host_path_matches_python_restatement(lib, oracle, k, alphabet, n_reads, rlen, tlen):
    """Randomised groups (low-complexity alphabets give repeats, branches and many equal-weight paths): C++ == Python for
    anchored paths, the auto mode, and under random node-id permutations (node_rank = the reference's boomphf order stand-in)."""
    rng = np.random.default_rng(k * 1000 + len(alphabet))
    n_groups = 24
    reads, goff = [], [0]
    templates = []
    for g in range(n_groups):
        t = "".join(rng.choice(list(alphabet), tlen))
        templates.append(t)
        for _ in range(n_reads):
            s = int(rng.integers(0, tlen - rlen + 1))
            r = list(t[s:s + rlen])
            for i in range(len(r)):
                if rng.random() < 0.02:
                    r[i] = str(rng.choice(list("ACGT")))
            reads.append("".join(r))
        goff.append(len(reads))
    data, offs, _ = oracle.strings_to_arrow(reads)
    goff = np.array(goff)
    n_found = 0
    for mc in (1, 2):
        b = oracle.count_prune_batch(data, offs, goff, k, mc, emit_csr=True)
        K = b.k_eff
        for g in range(n_groups):
            nodes = oracle.batch_group_as_tuples(b, g)
            if not nodes:
                continue
            t = templates[g]
            for (sa, ea) in ((t[:5], t[-5:]), (t[3:9], t[-12:-4]), ("", ""), (nodes[0][0][:3], nodes[-1][0][-2:])):
                want = oracle.path_finding_contig(nodes, K, sa, ea)
                got = _host_path(lib, b, g, sa, ea)
                assert got == want, (k, mc, g, sa, ea)
                n_found += want is not None
            assert _host_path(lib, b, g, None, None) == oracle.auto_path_finding_contig(nodes, K), (k, mc, g, "auto")
            # no rank: the container-free procedure the GPU runs per group (pathfind_core.h); identity rank: the petgraph restatement
            ident = np.arange(len(nodes))
            for (sa, ea) in ((t[:5], t[-5:]), ("", ""), (None, None)):
                assert _host_path(lib, b, g, sa, ea) == _host_path(lib, b, g, sa, ea, rank=ident), (k, mc, g, sa, ea, "core vs petgraph")
            # a permuted node order changes edge insertion order, anchor order and tie-breaks consistently on both sides
            perm = rng.permutation(len(nodes))          # rank[i] = id of ascending node i
            order = [None] * len(nodes)
            for i, pid in enumerate(perm):
                order[int(pid)] = nodes[i][0]
            sa, ea = t[:4], t[-4:]
            assert _host_path(lib, b, g, sa, ea, rank=perm) == oracle.path_finding_contig(nodes, K, sa, ea, order=order), (k, mc, g, "perm")
            assert _host_path(lib, b, g, None, None, rank=perm) == oracle.auto_path_finding_contig(nodes, K, order=order), (k, mc, g, "perm auto")
    assert n_found > 10
