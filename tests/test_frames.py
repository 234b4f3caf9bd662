"""The frame-level entry (rogtk_b200/frames.py): reads as the reference's converters leave them on disk (Parquet of
fastq_to_parquet, Arrow IPC of bam_to_arrow_ipc) grouped by a key column and assembled in one call -- what
df.group_by(key).agg(rogtk.assemble_sequences(...)) does with one plugin call per group (reference rogtk/__init__.py:104-156)."""
import numpy as np
import pytest


def _frame(rng, n_groups=12):
    import pyarrow as pa
    rows = []
    for g in range(n_groups):
        t = "".join(rng.choice(list("ACGT"), 160))
        for _ in range(int(rng.integers(3, 9))):
            s = int(rng.integers(0, 60))
            rows.append((f"umi{g:03d}" if g != 5 else None, t[s:s + 100]))
    rows.append(("umi_empty", None))
    perm = rng.permutation(len(rows))
    rows = [rows[i] for i in perm]
    return pa.table({"read_id": [f"r{i}" for i in range(len(rows))], "umi": [r[0] for r in rows], "r1_seq": [r[1] for r in rows],
                     "r1_qual": ["I" * 10] * len(rows)})


def test_group_reads_and_readers(tmp_path):
    import pyarrow as pa
    import pyarrow.ipc as ipc
    import pyarrow.parquet as pq
    from rogtk_b200 import frames
    t = _frame(np.random.default_rng(3))
    keys, lists = frames.group_reads(t, "umi", "r1_seq")
    seen = []
    for u in t.column("umi").to_pylist():
        if u not in seen:
            seen.append(u)
    assert keys.to_pylist() == seen                                            # order of first appearance, nulls a group of their own
    umi, seq = t.column("umi").to_pylist(), t.column("r1_seq").to_pylist()
    for u, got in zip(keys.to_pylist(), lists.to_pylist()):
        assert got == [s for x, s in zip(umi, seq) if x == u]                  # original row order inside a group
    pq.write_table(t, tmp_path / "reads.parquet")
    with ipc.new_file(str(tmp_path / "reads.arrow"), t.schema) as w:
        w.write_table(t)
    with ipc.new_stream(str(tmp_path / "reads.arrows"), t.schema) as w:
        w.write_table(t)
    for name in ("reads.parquet", "reads.arrow", "reads.arrows"):
        assert frames.read_reads(str(tmp_path / name), ["umi", "r1_seq"]).equals(t.select(["umi", "r1_seq"]))
    k0, l0 = frames.group_reads(t.slice(0, 0), "umi", "r1_seq")
    assert len(l0) == 0


@pytest.mark.gpu
def test_assemble_frame_matches_per_group_reference_calls(engine, oracle, tmp_path):
    import pyarrow.parquet as pq
    from rogtk_b200 import frames
    t = _frame(np.random.default_rng(4), n_groups=20)
    pq.write_table(t, tmp_path / "reads.parquet")
    table = frames.read_reads(str(tmp_path / "reads.parquet"))
    umi, seq = table.column("umi").to_pylist(), table.column("r1_seq").to_pylist()
    for method, kw in (("compression", {}), ("shortest_path_auto", {})):
        out = frames.assemble_frame(engine, table, "umi", "r1_seq", k=15, min_coverage=1, method=method, min_length=30, **kw)
        assert out.column_names == ["umi", "assembled_sequences", "flags"] and out.num_rows == len(set(umi))
        for u, got, fl in zip(out.column("umi").to_pylist(), out.column("assembled_sequences").to_pylist(), out.column("flags").to_pylist()):
            reads = [s for x, s in zip(umi, seq) if x == u and s is not None]     # null reads are dropped (src/expressions.rs:737-740)
            want, info = oracle.assemble_sequences_py(reads, 15, 1, method, min_length=30, only_largest=True)
            if info.get("tie") and method == "compression":
                assert fl & 1
                continue
            assert got == (want[0] if want else ""), (u, method)


def test_read_fasta_sequences(tmp_path):
    """The FASTA side of fracture_fasta (reference src/fracture.rs:161-186 reads records with bio::io::fasta): header lines dropped,
    wrapped lines joined, blank lines and a missing final newline tolerated, text before the first header ignored; case and
    non-ACGT records are left to the device (upper-casing and the whole-read drop happen there)."""
    from rogtk_b200.fracture import read_fasta_sequences
    fa = tmp_path / "x.fa"
    fa.write_bytes(b"junk before any header\n>r1 desc\nACGT\nacgt\n\n>r2\r\nTTTT\r\n>empty\n>r3\nACGTN\nGG")
    assert read_fasta_sequences(fa) == [b"ACGTacgt", b"TTTT", b"", b"ACGTNGG"]
    (tmp_path / "none.fa").write_bytes(b"")
    assert read_fasta_sequences(tmp_path / "none.fa") == []
