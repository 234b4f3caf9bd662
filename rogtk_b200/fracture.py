"""The non-Polars entries of the path: `fracture_sequences` and `fracture_fasta` (reference src/fracture.rs:468-564).

Same names, argument order, defaults and error types as the PyO3 functions: a bad method / missing anchors raises ValueError
(`AssemblyMethod::from_str` -> PyValueError, src/fracture.rs:484-485, 526-527), a failed assembly RuntimeError, and
`fracture_fasta` raises ValueError("No contigs found with length >= N") when nothing is long enough (src/fracture.rs:494-500).
`export_graphs` / `prefix` only drive the reference's DOT/CSV side files (src/graph_viz.rs) and are accepted and ignored.

Every call runs on the GPU through `rogtk_gpu_assemble_batch` with one group; there is no CPU fallback.  For many groups use
`Engine.assemble_sequences` (one call per frame) -- a call per group is launch-bound by construction.
"""
from __future__ import annotations

import threading

import numpy as np

from .engine import Engine, RogtkGpuError

_lock = threading.Lock()
_engines: dict[int, Engine] = {}


def default_engine(device: int = 0) -> Engine:
    """One lazily created context per device for the function-style entries (the context serialises its callers)."""
    with _lock:
        if device not in _engines:
            _engines[device] = Engine(device=device)
        return _engines[device]


def _column(sequences) -> tuple[np.ndarray, np.ndarray]:
    raw = [s.encode() if isinstance(s, str) else bytes(s) for s in sequences]
    offsets = np.zeros(len(raw) + 1, dtype=np.int64)
    if raw:
        offsets[1:] = np.cumsum([len(r) for r in raw])
    data = np.frombuffer(b"".join(raw) or b"\0", dtype=np.uint8).copy()
    return data, offsets


def _method_errors(fn):
    def wrapped(*a, **kw):
        try:
            return fn(*a, **kw)
        except RogtkGpuError as e:
            if "Invalid assembly method" in e.msg or "anchor" in e.msg.lower():
                raise ValueError(e.msg) from None
            raise RuntimeError(e.msg) from None
    wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
    return wrapped


@_method_errors
def fracture_sequences(sequences, k, min_coverage, min_length=200, method="compression", start_anchor=None, end_anchor=None,
                       export_graphs=None, only_largest=None, auto_k=None, prefix=None, *, engine: Engine | None = None) -> str:
    """Contigs of one set of reads joined by newlines, or the largest one (reference src/fracture.rs:506-564).
    The listing order of equally valid contigs is this library's (ascending first k-mer; the reference's is its MPHF order,
    DESIGN.md 6)."""
    eng = engine or default_engine()
    data, offsets = _column(sequences)
    out = eng.assemble_sequences(data, offsets, np.array([0, len(offsets) - 1], dtype=np.int64), k=k, min_coverage=min_coverage,
                                 method=method, start_anchor=start_anchor, end_anchor=end_anchor, min_length=min_length,
                                 only_largest=bool(only_largest), auto_k=bool(auto_k))
    return out.strings()[0]


def read_fasta_sequences(path) -> list[bytes]:
    """Records of a FASTA file as byte strings (lines of a record concatenated, header dropped).  Upper-casing and the
    ACGT-only rule (src/fracture.rs:161-186) are applied on the device like for any other read."""
    seqs, cur, seen = [], [], False
    with open(path, "rb") as fh:
        for line in fh:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if seen:
                    seqs.append(b"".join(cur))
                cur, seen = [], True
            elif seen:
                cur.append(line)
    if seen:
        seqs.append(b"".join(cur))
    return seqs


@_method_errors
def fracture_fasta(fasta_path, k, min_coverage, min_length=200, method="compression", start_anchor=None, end_anchor=None,
                   export_graphs=None, *, engine: Engine | None = None) -> str:
    """Largest contig (>= min_length) of the reads in a FASTA file (reference src/fracture.rs:468-503, assemble_fasta :283-321)."""
    eng = engine or default_engine()
    min_length = 200 if min_length is None else int(min_length)
    data, offsets = _column(read_fasta_sequences(fasta_path))
    out = eng.assemble_sequences(data, offsets, np.array([0, len(offsets) - 1], dtype=np.int64), k=k, min_coverage=min_coverage,
                                 method=method, start_anchor=start_anchor, end_anchor=end_anchor, min_length=min_length,
                                 only_largest=True, auto_k=False)
    s = out.strings()[0]
    if not s:
        raise ValueError(f"No contigs found with length >= {min_length}")
    return s
