"""Pins the golden fixtures against the REAL debruijn crate: compares the output of tools/dump_filter_kmers (Rust; needs cargo,
which the build image does not have) with tests/golden/appendix_b.json.

    for every case:   printf '%s\n' <reads, '-' for null> > reads.txt
                      cargo run --release --manifest-path tools/dump_filter_kmers/Cargo.toml -- <k> <min_cov> reads.txt > got.tsv
    python tests/golden/check_dump.py <case index> got.tsv      -> prints OK / the first differing row

    python tests/golden/check_dump.py --emit-inputs DIR         -> writes reads_<i>.txt + run_all.sh for every case
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    cases = json.load(open(os.path.join(HERE, "appendix_b.json")))["cases"]
    if sys.argv[1] == "--emit-inputs":
        d = sys.argv[2]
        os.makedirs(d, exist_ok=True)
        lines = ["#!/bin/sh", "set -e"]
        for i, c in enumerate(cases):
            with open(os.path.join(d, f"reads_{i}.txt"), "w") as f:
                for r in c["reads"]:
                    f.write(("-" if r is None else r) + "\n")
            lines.append(f"cargo run --release --manifest-path tools/dump_filter_kmers/Cargo.toml -- {c['k']} {c['min_coverage']} "
                         f"{d}/reads_{i}.txt > {d}/got_{i}.tsv && python tests/golden/check_dump.py {i} {d}/got_{i}.tsv")
        with open(os.path.join(d, "run_all.sh"), "w") as f:
            f.write("\n".join(lines) + "\n")
        print(f"{len(cases)} cases -> {d}/run_all.sh")
        return
    c = cases[int(sys.argv[1])]
    got = [ln.rstrip("\n").split("\t") for ln in open(sys.argv[2]) if ln.strip()]
    want = list(zip(c["kmers"], c["exts"], c["counts"]))
    if len(got) != len(want):
        print(f"DIFF case {c['name']} k={c['k']} min_cov={c['min_coverage']}: {len(got)} rows from the crate, {len(want)} in the fixture")
        sys.exit(1)
    for i, (g, w) in enumerate(zip(got, want)):
        if (g[0], int(g[1], 16), int(g[2])) != (w[0], int(w[1]), int(w[2])):
            print(f"DIFF case {c['name']} row {i}: crate {g[:3]} fixture {w}")
            sys.exit(1)
    print(f"OK case {c['name']} k={c['k']} min_cov={c['min_coverage']}: {len(got)} rows identical -> set pinned_by_reference true")


if __name__ == "__main__":
    main()
