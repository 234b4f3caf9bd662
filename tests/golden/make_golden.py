"""Generates tests/golden/appendix_b.json.

The reference (Rust, crates debruijn/boomphf not vendored, no Rust toolchain in this image) cannot be run here, so
these vectors are NOT outputs of the reference binary.  Inputs are the literals of the reference's own unit tests
(reference src/fracture.rs:577-583, :615-617, :685-686, :714-724); expected values are produced by the oracle and
must reproduce the SHA-256 digests that SURVEY.md Appendix B.1 derived independently from the same inputs.  The
contig-level expectations that the reference's tests DO state (the 44-bp answer in the comment at
src/fracture.rs:611, anchors contained :660-663, empty result for absent anchors :705-706) are recorded alongside.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import dbg_oracle as o  # noqa: E402

B1 = ["GAGACTGCATGGGCTGGTGGGCGTCCGTCTGC", "GGGCTGGTGGGCGTCCGTCTGCTTTAGTGAGGGT"]
B2 = [
    "ATGCATGCATGCTAGCTGATCGATCGTAGCTAGCTAGCTGATCGATCGTACGTACGTACGTAGCTACGTACGTACGTAGCTAGCTGATCGTAGCTACGTAGCTAGCTAGCTGATCGTACGTACGT",
    "GTAGCTAGCTAGCTGATCGATCGTACGTACGTACGTAGCTACGTACGTACGTAGCTAGCTGATCGTAGCTACGTAGCTAGCTAGCTGATCGTACGTACGTAGCTGATCGATCGTAGCTACGTACGT",
    "GTACGTACGTACGTAGCTACGTACGTACGTAGCTAGCTGATCGTAGCTACGTAGCTAGCTAGCTGATCGTACGTACGTAGCTGATCGATCGTAGCTACGTACGTACGTAGCTACGTACGTACGTAG",
    "TACGTACGTACGTAGCTAGCTGATCGTAGCTACGTAGCTAGCTAGCTGATCGTACGTACGTAGCTGATCGATCGTAGCTACGTACGTACGTAGCTACGTACGTACGTAGCTAGCTGATCGTAGCT",
]
B3 = ["AAAACCCCCAAAAA", "TTTTTGGGGGTTTT"]
B5 = [B1[0], B1[1].lower(), "ACGTNACGTACGTACGTACGTACGT", None, "", "ACG", "GGGCTGGTGGGCGTCCGTCTGCTTTAGTGAGGGT"]

SURVEY_DIGESTS = {
    ("B1", 13, 1): "99fe09bed6d9832905ddd7778a712a87188f82eaa2f1ffb7fc5e8e62447ab252",
    ("B2", 20, 1): "356ba229aeed1699cfe9079935a25884bbb76463dceb540f4d472deab8d99d62",
    ("B2", 20, 2): "cc58976609519d7551b56ea31f869116034f5723e3fa06ff51dd9cb71f023e9f",
    ("B2", 20, 3): "09ff4606bcdaa2e096d8beceb2dd2c2d6578338039b24712d2fa378ad625a17d",
    ("B3", 4, 1): "ba0da3c116e4ab78162ac4e68027195238a15599386eb90e9ca7f64c198cc4ad",
    ("B4", 4, 1): "e841be1ecfcad06aaaeda3cbe69fb3be38d43a4ab0bb971bc4586f712a3a6ad0",
}

CASES = [("B1", B1, 13, 1), ("B1", B1, 13, 2), ("B2", B2, 20, 1), ("B2", B2, 20, 2), ("B2", B2, 20, 3), ("B3", B3, 4, 1),
         ("B4", B1, 4, 1), ("B5", B5, 13, 1), ("B5", B5, 13, 0), ("B5", B5, 20, 2), ("B1", B1, 8, 1), ("B2", B2, 40, 1),
         ("B2", B2, 64, 2)]


def main():
    out = {"note": "oracle-derived; NOT reference outputs (see make_golden.py docstring)", "cases": []}
    for name, reads, k, mc in CASES:
        d, off, v = o.strings_to_arrow(reads)
        b = o.count_prune_batch(d, off, np.array([0, len(reads)]), k, mc, validity=v, emit_csr=True)
        rows = o.batch_group_as_tuples(b, 0)
        assert rows == o.count_prune_group_py(reads, k, mc), (name, k, mc)
        dig = o.digest(b, 0)
        if (name, k, mc) in SURVEY_DIGESTS:
            assert dig == SURVEY_DIGESTS[(name, k, mc)], (name, k, mc, dig)
        out["cases"].append({
            "name": name, "reads": reads, "k": k, "min_coverage": mc, "k_eff": b.k_eff,
            "n_observations": b.n_observations, "sha256": dig, "pinned_by_survey": (name, k, mc) in SURVEY_DIGESTS,
            # no output of the real crates exists for any case (no Rust toolchain here): tools/dump_filter_kmers + check_dump.py
            # flip this to true the day somebody runs them
            "pinned_by_reference": False,
            "stats": [int(x) for x in b.stats[0]],
            "kmers": [r[0] for r in rows], "exts": [r[1] for r in rows], "counts": [r[2] for r in rows],
            "csr_row_ptr": [int(x) for x in b.csr_row_ptr], "csr_col": [int(x) for x in b.csr_col],
        })
    out["reference_test_expectations"] = {
        "test_full_assembly_with_path_finding": {
            "source": "src/fracture.rs:611,621-679", "reads": B1, "k": 13, "min_coverage": 1,
            "start_anchor": "GAGACTGCATGG", "end_anchor": "TTTAGTGAGGGT",
            "expected_contig": "GAGACTGCATGGGCTGGTGGGCGTCCGTCTGCTTTAGTGAGGGT"},
        "test_assembly_with_invalid_anchors": {
            "source": "src/fracture.rs:681-707", "reads": B3, "k": 4, "min_coverage": 1,
            "start_anchor": "NONEXISTENT", "end_anchor": "ALSONOTHERE", "expected": []},
    }
    path = os.path.join(os.path.dirname(__file__), "appendix_b.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
