"""Frames either side of the path: reads as the reference's converters leave them on disk -- Parquet written by
`fastq_to_parquet` (columns read_id / r1_seq / r1_qual, reference src/single_fastq.rs:32-36) or Arrow IPC written by
`bam_to_arrow_ipc*` (name / chrom / ... / sequence / quality_scores, src/bam.rs:3203-3224) plus whatever key column the user derived
(a UMI) -- grouped by that key and handed to the batched entry in ONE call, Arrow buffers all the way: what
`df.group_by(key).agg(rogtk.assemble_sequences(pl.col(seq), ...))` does with one plugin call per group in the reference
(rogtk/__init__.py:104-156).  pyarrow only; the FASTQ / BAM parsers themselves are out of scope."""
from __future__ import annotations

import numpy as np


def read_reads(path: str, columns: list[str] | None = None):
    """A pyarrow Table from a Parquet file or an Arrow IPC file / stream (decided by the magic bytes)."""
    import pyarrow as pa
    import pyarrow.ipc as ipc
    import pyarrow.parquet as pq
    with open(path, "rb") as fh:
        magic = fh.read(6)
    if magic[:4] == b"PAR1":
        return pq.read_table(path, columns=columns)
    if magic == b"ARROW1":
        t = ipc.open_file(pa.memory_map(path)).read_all()
    else:
        t = ipc.open_stream(pa.memory_map(path)).read_all()
    return t.select(columns) if columns else t


def group_reads(table, key: str, seq: str):
    """(keys, List(String) array with the reads of every key, rows in their original order inside a group; null keys form a group
    of their own like in Polars).  Keys come out in order of first appearance."""
    import pyarrow as pa
    import pyarrow.compute as pc
    n = table.num_rows
    if n == 0:
        return table.column(key).combine_chunks() if table.num_columns else pa.array([]), pa.array([], type=pa.large_list(pa.large_string()))
    kcol = table.column(key).combine_chunks()
    enc = pc.dictionary_encode(kcol, null_encoding="encode")               # codes in order of first appearance
    codes = np.asarray(enc.indices.to_numpy(zero_copy_only=False), dtype=np.int64)
    order = np.argsort(codes, kind="stable")                                # stable: original row order inside a group
    counts = np.bincount(codes, minlength=len(enc.dictionary))
    goff = np.zeros(len(counts) + 1, dtype=np.int64)
    np.cumsum(counts, out=goff[1:])
    scol = table.column(seq).combine_chunks()
    if pa.types.is_string(scol.type) or (hasattr(pa.types, "is_string_view") and pa.types.is_string_view(scol.type)):
        scol = scol.cast(pa.large_string())   # (take() on a frame of more than 2 GiB of reads needs 64-bit offsets)
    child = scol.take(pa.array(order))
    lists = pa.LargeListArray.from_arrays(pa.array(goff), child)
    return enc.dictionary, lists


def assemble_frame(engine, table, key: str, seq: str, k: int = 10, min_coverage: int = 5, method: str = "shortest_path",
                   start_anchor: str | None = None, end_anchor: str | None = None, min_length: int | None = None, auto_k: bool = False):
    """group_by(key).agg(assemble_sequences(seq, ...)) for the whole frame in one call: a pyarrow Table {key, assembled_sequences,
    flags} with one row per key (defaults are the reference's, rogtk/__init__.py:104-116)."""
    import pyarrow as pa
    keys, lists = group_reads(table, key, seq)
    contigs, flags = engine.assemble_arrow(lists, k, min_coverage, min_length=min_length or 0, method=method, start_anchor=start_anchor,
                                           end_anchor=end_anchor, auto_k=auto_k)
    return pa.table({key: keys, "assembled_sequences": contigs, "flags": pa.array(np.asarray(flags, dtype=np.uint8))})
