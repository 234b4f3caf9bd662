"""rogtk_b200 -- B200-native per-UMI de Bruijn k-mer count / prune engine (drop-in for the filter_kmers call behind
rogtk's `assemble_sequences` Polars expression, reference src/fracture.rs:105-158)."""
from .engine import (ContigBatch, DeviceBatch, Engine, GraphBatch, RogtkGpuError, StringBatch, count_prune_multi, estimate_k, k_eff,  # noqa: F401
                     shard_of)
from . import _ffi as ffi  # noqa: F401
from . import frames, sharding, spectrum, synth  # noqa: F401
from .fracture import fracture_fasta, fracture_sequences  # noqa: F401

__all__ = ["Engine", "GraphBatch", "DeviceBatch", "ContigBatch", "StringBatch", "count_prune_multi", "fracture_sequences", "fracture_fasta", "RogtkGpuError", "k_eff", "estimate_k", "shard_of", "ffi", "synth", "sharding", "spectrum", "frames"]
