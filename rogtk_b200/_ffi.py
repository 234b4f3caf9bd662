"""ctypes binding of include/rogtk_gpu.h (the C ABI a Rust/PyO3 host would bind with `extern "C"`).

The product path has no CPU fallback: if librogtk_gpu.so is missing this module raises, and if no B200 is visible
``rogtk_gpu_create`` fails with ROGTK_ERR_CUDA and :class:`rogtk_b200.Engine` raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librogtk_gpu.so")

ROGTK_OK = 0
ROGTK_ERR_INVALID_ARG = -1
ROGTK_ERR_CUDA = -2
ROGTK_ERR_OOM = -3
ROGTK_ERR_UNSUPPORTED = -4
ROGTK_ERR_INTERNAL = -5

UTF8_I32, UTF8_I64, UTF8_VIEW = 0, 1, 2

EMIT_CSR = 1 << 0
EMIT_STATS = 1 << 1
EXACT_K = 1 << 2
INPUT_ON_DEVICE = 1 << 3
OUTPUT_ON_DEVICE = 1 << 4
FORCE_SPILL = 1 << 5
NO_MULTIPASS = 1 << 6
UNITIG_NODE_PARALLEL = 1 << 7
NO_PARTITION = 1 << 8

# every symbol include/rogtk_gpu.h declares (tests/test_abi.py checks the .so exports all of them)
SYMBOLS = [
    "rogtk_gpu_abi_version", "rogtk_gpu_device_count", "rogtk_gpu_create", "rogtk_gpu_destroy", "rogtk_gpu_trim", "rogtk_gpu_last_error",
    "rogtk_gpu_k_eff", "rogtk_gpu_estimate_k", "rogtk_gpu_shard_of", "rogtk_gpu_count_prune_batch", "rogtk_gpu_free_batch",
    "rogtk_gpu_synth_reads", "rogtk_gpu_synth_reads_from", "rogtk_gpu_synth_reads_var", "rogtk_gpu_device_alloc", "rogtk_gpu_device_free", "rogtk_gpu_pinned_alloc",
    "rogtk_gpu_pinned_free", "rogtk_gpu_memcpy_h2d", "rogtk_gpu_memcpy_d2h", "rogtk_gpu_launch_count", "rogtk_gpu_flush_l2", "rogtk_gpu_debug_phase_cycles", "rogtk_gpu_debug_cache_stats",
    "rogtk_gpu_spectrum_owner", "rogtk_gpu_spectrum_split", "rogtk_gpu_spectrum_merge", "rogtk_gpu_rethreshold",
    "rogtk_gpu_compress_unitigs", "rogtk_gpu_free_contigs",
    "rogtk_gpu_assemble_batch", "rogtk_gpu_free_strings", "rogtk_gpu_sweep_batch", "rogtk_gpu_free_sweep",
    "rogtk_gpu_optimize_batch", "rogtk_gpu_free_optimize", "rogtk_host_shortest_path", "rogtk_gpu_count_prune_multi", "rogtk_gpu_nccl_unique_id", "rogtk_gpu_nccl_init",
    "rogtk_gpu_nccl_finalize", "rogtk_gpu_spectrum_exchange", "rogtk_gpu_spectrum_merge_runs",
]
METHODS = {"compression": 0, "shortest_path": 1, "shortest_path_auto": 2}


class StrCol(ctypes.Structure):
    _fields_ = [
        ("layout", ctypes.c_int32),
        ("n_buffers", ctypes.c_int32),
        ("n_rows", ctypes.c_int64),
        ("validity", ctypes.c_void_p),
        ("validity_bit_offset", ctypes.c_int64),
        ("offsets_or_views", ctypes.c_void_p),
        ("data", ctypes.POINTER(ctypes.c_void_p)),
        ("data_len", ctypes.POINTER(ctypes.c_int64)),
    ]


class GraphBatchC(ctypes.Structure):
    _fields_ = [
        ("n_groups", ctypes.c_int64),
        ("n_kmers", ctypes.c_int64),
        ("n_edges", ctypes.c_int64),
        ("n_observations", ctypes.c_int64),
        ("k_eff", ctypes.c_int32),
        ("on_device", ctypes.c_int32),
        ("kmer_offsets", ctypes.c_void_p),
        ("kmers_lo", ctypes.c_void_p),
        ("kmers_hi", ctypes.c_void_p),
        ("exts", ctypes.c_void_p),
        ("counts", ctypes.c_void_p),
        ("csr_row_ptr", ctypes.c_void_p),
        ("csr_col", ctypes.c_void_p),
        ("stats", ctypes.c_void_p),
        ("ms_total", ctypes.c_float),
        ("ms_count", ctypes.c_float),
        ("ms_hot_kernel", ctypes.c_float),
        ("n_multipass_groups", ctypes.c_int32),
        ("n_spilled_groups", ctypes.c_int64),
        ("ms_sort", ctypes.c_float),
        ("ms_csr", ctypes.c_float),
        ("ms_stats", ctypes.c_float),
        ("_pad1", ctypes.c_float),
        ("_owner", ctypes.c_void_p),
    ]


class ContigBatchC(ctypes.Structure):
    _fields_ = [
        ("n_groups", ctypes.c_int64),
        ("n_unitigs", ctypes.c_int64),
        ("n_bases", ctypes.c_int64),
        ("k_eff", ctypes.c_int32),
        ("on_device", ctypes.c_int32),
        ("unitig_offsets", ctypes.c_void_p),
        ("seq_offsets", ctypes.c_void_p),
        ("seq", ctypes.c_void_p),
        ("counts", ctypes.c_void_p),
        ("exts", ctypes.c_void_p),
        ("cyclic", ctypes.c_void_p),
        ("largest", ctypes.c_void_p),
        ("largest_flags", ctypes.c_void_p),
        ("ms_total", ctypes.c_float),
        ("_pad0", ctypes.c_float),
        ("_owner", ctypes.c_void_p),
    ]


class AssemblyParamsC(ctypes.Structure):
    _fields_ = [
        ("k", ctypes.c_int64),
        ("min_coverage", ctypes.c_uint32),
        ("method", ctypes.c_int32),
        ("start_anchor", ctypes.c_char_p),
        ("end_anchor", ctypes.c_char_p),
        ("min_length", ctypes.c_int64),
        ("only_largest", ctypes.c_int32),
        ("auto_k", ctypes.c_int32),
        ("start_anchor_offsets", ctypes.c_void_p),
        ("start_anchor_data", ctypes.c_void_p),
        ("end_anchor_offsets", ctypes.c_void_p),
        ("end_anchor_data", ctypes.c_void_p),
        ("node_rank", ctypes.c_void_p),
        ("host_threads", ctypes.c_int32),
        ("_pad", ctypes.c_int32),
    ]


class StringBatchC(ctypes.Structure):
    _fields_ = [
        ("n_rows", ctypes.c_int64),
        ("n_bytes", ctypes.c_int64),
        ("offsets", ctypes.c_void_p),
        ("data", ctypes.c_void_p),
        ("flags", ctypes.c_void_p),
        ("ms_device", ctypes.c_float),
        ("ms_host", ctypes.c_float),
        ("n_count_passes", ctypes.c_int64),
        ("_owner", ctypes.c_void_p),
    ]


class SweepParamsC(ctypes.Structure):
    _fields_ = [
        ("k_start", ctypes.c_int64), ("k_end", ctypes.c_int64), ("k_step", ctypes.c_int64),
        ("cov_start", ctypes.c_int64), ("cov_end", ctypes.c_int64), ("cov_step", ctypes.c_int64),
        ("method", ctypes.c_int32),
        ("host_threads", ctypes.c_int32),
        ("start_anchor", ctypes.c_char_p),
        ("end_anchor", ctypes.c_char_p),
    ]


class SweepResultC(ctypes.Structure):
    _fields_ = [
        ("n_groups", ctypes.c_int64), ("n_cells", ctypes.c_int64),
        ("k", ctypes.c_void_p), ("min_coverage", ctypes.c_void_p), ("contig_length", ctypes.c_void_p),
        ("ms_device", ctypes.c_float), ("ms_host", ctypes.c_float),
        ("n_count_passes", ctypes.c_int64),
        ("_owner", ctypes.c_void_p),
    ]


class OptimizeParamsC(ctypes.Structure):
    _fields_ = [
        ("start_k", ctypes.c_int64), ("start_min_coverage", ctypes.c_int64), ("max_iterations", ctypes.c_int64),
        ("explore_k", ctypes.c_int32), ("prioritize_length", ctypes.c_int32),
        ("method", ctypes.c_int32), ("host_threads", ctypes.c_int32),
        ("start_anchor", ctypes.c_char_p), ("end_anchor", ctypes.c_char_p),
    ]


class OptimizeResultC(ctypes.Structure):
    _fields_ = [
        ("contig", StringBatchC),
        ("k", ctypes.c_void_p), ("min_coverage", ctypes.c_void_p), ("length", ctypes.c_void_p), ("input_sequences", ctypes.c_void_p),
        ("n_cells_evaluated", ctypes.c_int64),
        ("n_probes", ctypes.c_int64),
        ("_owner", ctypes.c_void_p),
    ]


_lib = None


def load() -> ctypes.CDLL:
    """Load librogtk_gpu.so (built in-tree by rogtk_b200/csrc/Makefile).  Raises if it is missing: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make -C rogtk_b200/csrc` (or __graft_entry__.build()); "
            "rogtk_b200 has no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i64, i32, u32, u64, sz = (ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint64,
                                  ctypes.c_size_t)
    lib.rogtk_gpu_abi_version.restype = ctypes.c_int
    lib.rogtk_gpu_device_count.restype = ctypes.c_int
    lib.rogtk_gpu_create.argtypes = [ctypes.c_int, sz, ctypes.POINTER(ctypes.c_int)]
    lib.rogtk_gpu_create.restype = vp
    lib.rogtk_gpu_destroy.argtypes = [vp]
    lib.rogtk_gpu_trim.argtypes = [vp]
    lib.rogtk_gpu_trim.restype = ctypes.c_int
    lib.rogtk_gpu_last_error.argtypes = [vp]
    lib.rogtk_gpu_last_error.restype = ctypes.c_char_p
    lib.rogtk_gpu_k_eff.argtypes = [i64]
    lib.rogtk_gpu_k_eff.restype = ctypes.c_int
    lib.rogtk_gpu_estimate_k.argtypes = [vp, vp, i64]
    lib.rogtk_gpu_estimate_k.restype = i64
    lib.rogtk_gpu_shard_of.argtypes = [u64, i32]
    lib.rogtk_gpu_shard_of.restype = i32
    lib.rogtk_gpu_count_prune_batch.argtypes = [vp, ctypes.POINTER(StrCol), vp, i64, i64, u32, u32,
                                                ctypes.POINTER(GraphBatchC)]
    lib.rogtk_gpu_count_prune_batch.restype = ctypes.c_int
    lib.rogtk_gpu_free_batch.argtypes = [vp, ctypes.POINTER(GraphBatchC)]
    lib.rogtk_gpu_synth_reads.argtypes = [vp, u64, i64, i64, i32, i32, i32, u32, vp, vp, vp]
    lib.rogtk_gpu_synth_reads.restype = ctypes.c_int
    lib.rogtk_gpu_synth_reads_from.argtypes = [vp, u64, i64, i64, i32, i32, i32, i32, u32, vp, vp, vp, vp]
    lib.rogtk_gpu_synth_reads_from.restype = ctypes.c_int
    lib.rogtk_gpu_synth_reads_var.argtypes = [vp, u64, i64, i64, vp, i64, i32, i32, u32, vp, vp]
    lib.rogtk_gpu_synth_reads_var.restype = ctypes.c_int
    lib.rogtk_gpu_device_alloc.argtypes = [vp, sz]
    lib.rogtk_gpu_device_alloc.restype = vp
    lib.rogtk_gpu_device_free.argtypes = [vp, vp]
    lib.rogtk_gpu_pinned_alloc.argtypes = [vp, sz]
    lib.rogtk_gpu_pinned_alloc.restype = vp
    lib.rogtk_gpu_pinned_free.argtypes = [vp, vp]
    lib.rogtk_gpu_memcpy_h2d.argtypes = [vp, vp, vp, sz]
    lib.rogtk_gpu_memcpy_h2d.restype = ctypes.c_int
    lib.rogtk_gpu_memcpy_d2h.argtypes = [vp, vp, vp, sz]
    lib.rogtk_gpu_memcpy_d2h.restype = ctypes.c_int
    lib.rogtk_gpu_launch_count.argtypes = [vp]
    lib.rogtk_gpu_launch_count.restype = i64
    lib.rogtk_gpu_debug_phase_cycles.argtypes = [vp, ctypes.c_int, vp]
    lib.rogtk_gpu_debug_phase_cycles.restype = ctypes.c_int
    lib.rogtk_gpu_debug_cache_stats.argtypes = [vp, vp]
    lib.rogtk_gpu_debug_cache_stats.restype = ctypes.c_int
    lib.rogtk_gpu_spectrum_owner.argtypes = [u64, u64, ctypes.c_int, ctypes.c_int]
    lib.rogtk_gpu_spectrum_owner.restype = ctypes.c_int
    lib.rogtk_gpu_spectrum_split.argtypes = [vp, vp, vp, i64, ctypes.c_int, ctypes.c_int, vp]
    lib.rogtk_gpu_spectrum_split.restype = ctypes.c_int
    lib.rogtk_gpu_spectrum_merge.argtypes = [vp, vp, vp, vp, vp, i64, ctypes.c_int, u32, u32, ctypes.POINTER(GraphBatchC)]
    lib.rogtk_gpu_spectrum_merge.restype = ctypes.c_int
    lib.rogtk_gpu_rethreshold.argtypes = [vp, ctypes.POINTER(GraphBatchC), u32, u32, ctypes.POINTER(GraphBatchC)]
    lib.rogtk_gpu_rethreshold.restype = ctypes.c_int
    lib.rogtk_gpu_compress_unitigs.argtypes = [vp, ctypes.POINTER(GraphBatchC), i64, u32, ctypes.POINTER(ContigBatchC)]
    lib.rogtk_gpu_compress_unitigs.restype = ctypes.c_int
    lib.rogtk_gpu_free_contigs.argtypes = [vp, ctypes.POINTER(ContigBatchC)]
    lib.rogtk_gpu_assemble_batch.argtypes = [vp, ctypes.POINTER(StrCol), vp, i64, ctypes.POINTER(AssemblyParamsC), u32,
                                             ctypes.POINTER(StringBatchC)]
    lib.rogtk_gpu_assemble_batch.restype = ctypes.c_int
    lib.rogtk_gpu_free_strings.argtypes = [vp, ctypes.POINTER(StringBatchC)]
    lib.rogtk_gpu_sweep_batch.argtypes = [vp, ctypes.POINTER(StrCol), vp, i64, ctypes.POINTER(SweepParamsC), u32, ctypes.POINTER(SweepResultC)]
    lib.rogtk_gpu_sweep_batch.restype = ctypes.c_int
    lib.rogtk_gpu_free_sweep.argtypes = [vp, ctypes.POINTER(SweepResultC)]
    lib.rogtk_gpu_optimize_batch.argtypes = [vp, ctypes.POINTER(StrCol), vp, i64, ctypes.POINTER(OptimizeParamsC), u32,
                                             ctypes.POINTER(OptimizeResultC)]
    lib.rogtk_gpu_optimize_batch.restype = ctypes.c_int
    lib.rogtk_gpu_free_optimize.argtypes = [vp, ctypes.POINTER(OptimizeResultC)]
    lib.rogtk_host_shortest_path.argtypes = [ctypes.c_int, i32, vp, vp, vp, vp, vp, vp, ctypes.c_char_p, ctypes.c_char_p, vp, i64]
    lib.rogtk_host_shortest_path.restype = i64
    lib.rogtk_gpu_count_prune_multi.argtypes = [ctypes.POINTER(vp), ctypes.c_int, ctypes.POINTER(StrCol), vp, i64, vp, i64, u32, u32,
                                                ctypes.POINTER(GraphBatchC), vp]
    lib.rogtk_gpu_count_prune_multi.restype = ctypes.c_int
    lib.rogtk_gpu_nccl_unique_id.argtypes = [vp]
    lib.rogtk_gpu_nccl_unique_id.restype = ctypes.c_int
    lib.rogtk_gpu_nccl_init.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int]
    lib.rogtk_gpu_nccl_init.restype = ctypes.c_int
    lib.rogtk_gpu_nccl_finalize.argtypes = [vp]
    lib.rogtk_gpu_nccl_finalize.restype = ctypes.c_int
    lib.rogtk_gpu_spectrum_exchange.argtypes = [vp, ctypes.POINTER(GraphBatchC), ctypes.POINTER(GraphBatchC), u32, u32, ctypes.POINTER(GraphBatchC), vp]
    lib.rogtk_gpu_spectrum_exchange.restype = ctypes.c_int
    lib.rogtk_gpu_spectrum_merge_runs.argtypes = [vp, vp, ctypes.c_int, u32, u32, ctypes.POINTER(GraphBatchC)]
    lib.rogtk_gpu_spectrum_merge_runs.restype = ctypes.c_int
    lib.rogtk_gpu_flush_l2.argtypes = [vp]
    lib.rogtk_gpu_flush_l2.restype = ctypes.c_int
    _lib = lib
    return lib
