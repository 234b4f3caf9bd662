"""Which readings of the un-vendored k-mer engine satisfy the reference's own unit tests?  (CPU only; TEST INFRASTRUCTURE.)

The arithmetic of the hot path lives in debruijn ^0.3.4 / boomphf ^0.5, which are not in /root/reference and cannot be built
here, so the oracle is pinned only by the crates' published semantics.  This tool runs the inputs of the four unit
tests the reference holds for this path (src/fracture.rs:593-761) through every alternative reading of those semantics and
tabulates which readings let each assertion hold:

  strand      stranded (filter_kmers(.., stranded = true, ..) as called at src/fracture.rs:114)  |  RC-canonical (k-mer = min(k-mer,
              reverse complement), exts flipped with it; what `stranded = false` would do)
  k           K_eff rounding (Kmer4/8/16/32/64 type constant, src/fracture.rs:246-256)  |  the user's k itself
  exts        dangling (survivors keep neighbour bits of pruned k-mers)  |  censored (bits of pruned k-mers dropped
              before graph construction) -- only matters when min_coverage prunes something

Assertions (reference src/fracture.rs):
  T1 :604-606  Compression, k=20, min_cov=1 on the four 125-bp reads: contigs non-empty and contigs.first().len() > 150.
               The listing order is boomphf's, so a reading "can satisfy" T1 iff SOME contig is longer than 150 bp.
  T2 :660-663  ShortestPath, k=13, min_cov=1, anchors GAGACTGCATGG / TTTAGTGAGGGT: contig contains both anchors.
  T3 :705-706  ShortestPath, k=4, anchors absent: no contigs.
  T4 :757-760  k=4: Compression contigs non-empty AND ShortestPath (12-base anchors) contigs non-empty.

usage: python tools/parity_sensitivity.py            (prints a markdown table; DESIGN.md 3 carries a copy)
"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dbg_oracle as o  # noqa: E402

T1_READS = [
    "ATGCATGCATGCTAGCTGATCGATCGTAGCTAGCTAGCTGATCGATCGTACGTACGTACGTAGCTACGTACGTACGTAGCTAGCTGATCGTAGCTACGTAGCTAGCTAGCTGATCGTACGTACGT",
    "GTAGCTAGCTAGCTGATCGATCGTACGTACGTACGTAGCTACGTACGTACGTAGCTAGCTGATCGTAGCTACGTAGCTAGCTAGCTGATCGTACGTACGTAGCTGATCGATCGTAGCTACGTACGT",
    "GTACGTACGTACGTAGCTACGTACGTACGTAGCTAGCTGATCGTAGCTACGTAGCTAGCTAGCTGATCGTACGTACGTAGCTGATCGATCGTAGCTACGTACGTACGTAGCTACGTACGTACGTAG",
    "TACGTACGTACGTAGCTAGCTGATCGTAGCTACGTAGCTAGCTAGCTGATCGTACGTACGTAGCTGATCGATCGTAGCTACGTACGTACGTAGCTACGTACGTACGTAGCTAGCTGATCGTAGCT",
]
T2_READS = ["GAGACTGCATGGGCTGGTGGGCGTCCGTCTGC", "GGGCTGGTGGGCGTCCGTCTGCTTTAGTGAGGGT"]
T3_READS = ["AAAACCCCCAAAAA", "TTTTTGGGGGTTTT"]
COMP = str.maketrans("ACGT", "TGCA")


def rc(s):
    return s.translate(COMP)[::-1]


def flip_exts(e):
    """Exts::rc(): left <-> right with complemented bases"""
    out = 0
    for b in range(4):
        if e & (1 << b):
            out |= 16 << (3 - b)
        if e & (16 << b):
            out |= 1 << (3 - b)
    return out


def nodes_under(reads, k, min_cov, strand, kmode, exts_mode):
    K = o.k_eff_py(k) if kmode == "K_eff" else k
    if K < 0:
        return [], K
    nodes = o.count_prune_group_py(reads, K, 0, exact_k=True)          # everything, then prune here
    if strand == "canonical":
        tab = {}
        for km, e, c in nodes:
            r = rc(km)
            if r < km:
                km, e = r, flip_exts(e)
            ent = tab.setdefault(km, [0, 0])
            ent[0] = min(65535, ent[0] + c)
            ent[1] |= e
        nodes = [(km, ent[1], ent[0]) for km, ent in sorted(tab.items())]
    nodes = [(km, e, c) for km, e, c in nodes if c >= min_cov]
    if exts_mode == "censored":
        present = {km for km, _, _ in nodes} | ({rc(km) for km, _, _ in nodes} if strand == "canonical" else set())
        out = []
        for km, e, c in nodes:
            f = 0
            for b in range(4):
                if e & (1 << b) and ("ACGT"[b] + km[:-1]) in present:
                    f |= 1 << b
                if e & (16 << b) and (km[1:] + "ACGT"[b]) in present:
                    f |= 16 << b
            out.append((km, f, c))
        nodes = out
    return nodes, K


def doubled(nodes):
    """orientation-doubled node set of a canonical graph: every node also as its reverse complement with flipped exts"""
    d = {km: (e, c) for km, e, c in nodes}
    for km, e, c in nodes:
        r = rc(km)
        if r not in d:
            d[r] = (flip_exts(e), c)
    return [(km, e, c) for km, (e, c) in sorted(d.items())]


def contigs_under(nodes, K, strand):
    if strand == "stranded":
        return sorted(o.compress_unitigs(nodes, K), key=len, reverse=True)
    # canonical: maximal non-branching chains in the doubled graph; a chain and its mirror image are one contig
    cs = o.compress_unitigs(doubled(nodes), K)
    seen, out = set(), []
    for c in cs:
        key = min(c, rc(c))
        if key not in seen:
            seen.add(key)
            out.append(key)
    return sorted(out, key=len, reverse=True)


def path_under(nodes, K, strand, sa, ea):
    ns = nodes if strand == "stranded" else doubled(nodes)
    return o.path_finding_contig(ns, K, sa, ea)


def main():
    rows = []
    for strand, kmode, exts_mode in itertools.product(("stranded", "canonical"), ("K_eff", "true k"), ("dangling", "censored")):
        n1, K1 = nodes_under(T1_READS, 20, 1, strand, kmode, exts_mode)
        c1 = contigs_under(n1, K1, strand)
        t1 = bool(c1) and len(c1[0]) > 150
        n2, K2 = nodes_under(T2_READS, 13, 1, strand, kmode, exts_mode)
        p2 = path_under(n2, K2, strand, "GAGACTGCATGG", "TTTAGTGAGGGT")
        t2 = p2 is not None and "GAGACTGCATGG" in p2 and "TTTAGTGAGGGT" in p2
        n3, K3 = nodes_under(T3_READS, 4, 1, strand, kmode, exts_mode)
        t3 = path_under(n3, K3, strand, "NONEXISTENT", "ALSONOTHERE") is None
        n4, K4 = nodes_under(T2_READS, 4, 1, strand, kmode, exts_mode)
        c4 = contigs_under(n4, K4, strand)
        p4 = path_under(n4, K4, strand, "GAGACTGCATGG", "TTTAGTGAGGGT")
        t4 = bool(c4) and p4 is not None
        rows.append((strand, kmode, exts_mode, f"K={K1}: {len(n1)} nodes, contigs {[len(c) for c in c1[:4]]}", t1,
                     f"K={K2}: {len(p2) if p2 else 0} bp", t2, t3, f"{len(c4)} contigs, path {'found' if p4 else 'none'}", t4))
    print("| strand | k | exts | T1 facts | T1 | T2 facts | T2 | T3 | T4 facts | T4 |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        print("| " + " | ".join("yes" if x is True else "NO" if x is False else str(x) for x in r) + " |")
    ok = [r for r in rows if r[4] and r[6] and r[7]]
    print(f"\nreadings satisfying T1, T2 and T3 together: {len(ok)} of {len(rows)}"
          + ("" if not ok else ": " + "; ".join("/".join(r[:3]) for r in ok)))
    print("T4 cannot hold under any reading: its anchors (12 bases) are longer than the K = 4 node strings they are matched against "
          "with starts_with / ends_with (src/djfind.rs:133-140), so path finding returns no contig.")


if __name__ == "__main__":
    main()
