"""Host-side driver of the GPU k-mer engine (thin layer over the C ABI, include/rogtk_gpu.h).

`Engine.count_prune` is the batched form of the reference's per-group
``filter_kmers::<K,..>(seq_tuples, &Box::new(CountFilter::new(min_coverage)), true, true, 4)`` call
(reference src/fracture.rs:110-116): one call carries every group (UMI) of a frame.  Inputs are Arrow string buffers
(data + offsets [+ validity]) plus group row offsets, exactly what the Polars plugin receives
(reference src/expressions.rs:734-740); outputs are the three vectors the reference hands to ``BoomHashMap2::new``
(ascending k-mers, Exts bytes, saturating u16 counts) concatenated over groups.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import _ffi

_ERR = {-1: "invalid argument", -2: "CUDA error", -3: "out of memory", -4: "unsupported", -5: "internal error"}


class RogtkGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rogtk_gpu error {code} ({_ERR.get(code, '?')}): {msg}")
        self.code = code
        self.msg = msg


@dataclass
class GraphBatch:
    """Pruned de Bruijn nodes of every group.  Survivors of group g are [kmer_offsets[g], kmer_offsets[g+1])."""
    k_eff: int
    n_observations: int
    kmer_offsets: np.ndarray          # int64 [n_groups+1]
    kmers_lo: np.ndarray              # uint64, ascending inside a group
    kmers_hi: np.ndarray | None       # uint64 when k_eff > 32
    exts: np.ndarray                  # uint8  (low nibble left, high nibble right, bit = base code)
    counts: np.ndarray                # uint16 saturating
    stats: np.ndarray | None = None   # uint32 [n_groups,3]: node, terminal, isolated (fracture.rs:124-141)
    csr_row_ptr: np.ndarray | None = None
    csr_col: np.ndarray | None = None
    ms_total: float = 0.0
    ms_count: float = 0.0
    ms_hot_kernel: float = 0.0
    n_spilled_groups: int = 0
    n_multipass_groups: int = 0

    @property
    def n_groups(self) -> int:
        return len(self.kmer_offsets) - 1

    def group(self, g: int):
        a, b = int(self.kmer_offsets[g]), int(self.kmer_offsets[g + 1])
        hi = self.kmers_hi[a:b] if self.kmers_hi is not None else None
        return self.kmers_lo[a:b], hi, self.exts[a:b], self.counts[a:b]


@dataclass
class DeviceBatch:
    """Result left in HBM (ROGTK_OUTPUT_ON_DEVICE); only sizes and timings are visible on the host."""
    k_eff: int
    n_groups: int
    n_kmers: int
    n_edges: int
    n_observations: int
    ms_total: float
    ms_count: float
    ms_hot_kernel: float
    n_spilled_groups: int
    n_multipass_groups: int
    _c: _ffi.GraphBatchC
    _engine: "Engine"

    def free(self):
        if self._c is not None:
            self._engine._lib.rogtk_gpu_free_batch(self._engine._ctx, ctypes.byref(self._c))
            self._c = None

    @property
    def stage_ms(self) -> dict:
        """device time per stage of the pipeline that produced this batch (CUDA events on the library's stream)"""
        c = self._c
        return {"hot_kernel": float(c.ms_hot_kernel), "count": float(c.ms_count), "sort": float(c.ms_sort), "csr": float(c.ms_csr),
                "stats": float(c.ms_stats), "total": float(c.ms_total)}

    def to_host(self) -> GraphBatch:
        e, c = self._engine, self._c

        def pull(ptr, n, dt):
            out = np.empty(n, dtype=dt)
            if n and ptr:
                e._check(e._lib.rogtk_gpu_memcpy_d2h(e._ctx, out.ctypes.data, ptr, out.nbytes))
            return out

        return GraphBatch(
            k_eff=self.k_eff, n_observations=self.n_observations,
            kmer_offsets=pull(c.kmer_offsets, self.n_groups + 1, np.int64),
            kmers_lo=pull(c.kmers_lo, self.n_kmers, np.uint64),
            kmers_hi=pull(c.kmers_hi, self.n_kmers, np.uint64) if c.kmers_hi else None,
            exts=pull(c.exts, self.n_kmers, np.uint8), counts=pull(c.counts, self.n_kmers, np.uint16),
            stats=pull(c.stats, 3 * self.n_groups, np.uint32).reshape(-1, 3) if c.stats else None,
            csr_row_ptr=pull(c.csr_row_ptr, self.n_kmers + 1, np.int64) if c.csr_row_ptr else None,
            csr_col=pull(c.csr_col, self.n_edges, np.int32) if c.csr_col else None,
            ms_total=self.ms_total, ms_count=self.ms_count, ms_hot_kernel=self.ms_hot_kernel,
            n_spilled_groups=self.n_spilled_groups, n_multipass_groups=self.n_multipass_groups)


@dataclass
class ContigBatch:
    """Unitigs of every group (reference src/fracture.rs:351-383) and the plugin's per-group answer
    (min_length filter + longest contig, src/fracture.rs:258-277).  Unitigs of a group are listed by ascending first k-mer."""
    k_eff: int
    unitig_offsets: np.ndarray   # int64 [n_groups+1]
    seq_offsets: np.ndarray      # int64 [n_unitigs+1]
    seq: np.ndarray              # uint8 ASCII
    counts: np.ndarray           # uint16 saturating sum of node counts
    exts: np.ndarray             # uint8
    cyclic: np.ndarray           # uint8
    largest: np.ndarray          # int64 [n_groups], -1 = no contig of at least min_length
    largest_flags: np.ndarray    # uint8 [n_groups]: bit 0 tie in length, bit 1 cyclic
    ms_total: float = 0.0

    @property
    def n_groups(self) -> int:
        return len(self.unitig_offsets) - 1

    def unitig(self, u: int) -> str:
        return self.seq[int(self.seq_offsets[u]):int(self.seq_offsets[u + 1])].tobytes().decode("ascii")

    def group(self, g: int):
        """[(sequence, count, exts, cyclic)] of group g."""
        a, b = int(self.unitig_offsets[g]), int(self.unitig_offsets[g + 1])
        return [(self.unitig(u), int(self.counts[u]), int(self.exts[u]), bool(self.cyclic[u])) for u in range(a, b)]

    def assembled(self, g: int):
        """What assemble_sequences_expr returns for the group (one contig, '' if none), and the ambiguity flags."""
        u = int(self.largest[g])
        return (self.unitig(u) if u >= 0 else ""), int(self.largest_flags[g])


def k_eff(k: int) -> int:
    """reference src/fracture.rs:246-256: K = 4/8/16/32/64, the smallest >= k; -1 when k > 64."""
    return int(_ffi.load().rogtk_gpu_k_eff(int(k)))


def estimate_k(offsets: np.ndarray, validity: np.ndarray | None = None) -> int:
    """reference src/fracture.rs:24-54 (auto_k)."""
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    vp = np.ascontiguousarray(validity, dtype=np.uint8).ctypes.data if validity is not None else None
    return int(_ffi.load().rogtk_gpu_estimate_k(offsets.ctypes.data, vp, len(offsets) - 1))


def shard_of(umi_id: int, n_shards: int) -> int:
    """UMI -> GPU: splitmix64(umi_id) mod n_shards (no collective on the per-UMI path)."""
    return int(_ffi.load().rogtk_gpu_shard_of(int(umi_id) & 0xFFFFFFFFFFFFFFFF, int(n_shards)))


@dataclass
class StringBatch:
    """One string per group (Arrow LargeUtf8 buffers).  flags: bit 0 several longest contigs, bit 1 cyclic contig, bit 2 the group
    used its own k (auto_k), bit 3 several contigs listed (ascending first k-mer; the reference lists them in boomphf order)."""
    offsets: np.ndarray
    data: np.ndarray
    flags: np.ndarray
    ms_device: float = 0.0
    ms_host: float = 0.0
    n_count_passes: int = 0

    def strings(self) -> list[str]:
        b = self.data.tobytes()
        return [b[int(self.offsets[i]):int(self.offsets[i + 1])].decode() for i in range(len(self.offsets) - 1)]

    def to_arrow(self):
        import pyarrow as pa
        return pa.LargeStringArray.from_buffers(len(self.offsets) - 1, pa.py_buffer(self.offsets), pa.py_buffer(self.data))


class Engine:
    """One context == one GPU + one stream.  Not a fallback-capable object: construction fails without a B200."""

    def __init__(self, device: int = 0, workspace_bytes: int = 0):
        self._lib = _ffi.load()
        err = ctypes.c_int(0)
        self._ctx = self._lib.rogtk_gpu_create(int(device), int(workspace_bytes), ctypes.byref(err))
        if not self._ctx:
            raise RogtkGpuError(err.value, "rogtk_gpu_create failed (no usable sm_100 GPU?) - there is no CPU fallback")
        self.device = device

    def trim(self):
        """Return the context's cached device blocks and page-locked buffers to the driver (rogtk_gpu_trim)."""
        self._check(self._lib.rogtk_gpu_trim(self._ctx))

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.rogtk_gpu_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise RogtkGpuError(rc, self._lib.rogtk_gpu_last_error(self._ctx).decode())

    @property
    def launch_count(self) -> int:
        return int(self._lib.rogtk_gpu_launch_count(self._ctx))

    def phase_cycles(self, enable: bool = True) -> np.ndarray:
        """diagnostics: per-phase SM cycles of the last batch inside k_smem_count (summed over CTAs)"""
        out = np.zeros(16, dtype=np.uint64)
        self._check(self._lib.rogtk_gpu_debug_phase_cycles(self._ctx, int(enable), out.ctypes.data))
        return out

    def cache_stats(self) -> dict:
        """Counters of the context's device-memory cache (rogtk_gpu_debug_cache_stats)."""
        out = np.zeros(7, dtype=np.int64)
        self._check(self._lib.rogtk_gpu_debug_cache_stats(self._ctx, out.ctypes.data))
        return dict(zip(("n_malloc", "n_reuse", "idle_bytes", "live_bytes", "peak_bytes", "idle_blocks", "n_released_for_others"), map(int, out)))

    def flush_l2(self):
        self._check(self._lib.rogtk_gpu_flush_l2(self._ctx))

    # ---- raw memory helpers (harness use) ------------------------------------------------------------------------
    def device_alloc(self, nbytes: int) -> int:
        p = self._lib.rogtk_gpu_device_alloc(self._ctx, int(nbytes))
        if not p:
            raise RogtkGpuError(-3, self._lib.rogtk_gpu_last_error(self._ctx).decode())
        return int(p)

    def device_free(self, p: int):
        self._lib.rogtk_gpu_device_free(self._ctx, ctypes.c_void_p(p))

    def pinned_alloc(self, nbytes: int) -> int:
        p = self._lib.rogtk_gpu_pinned_alloc(self._ctx, int(nbytes))
        if not p:
            raise RogtkGpuError(-3, self._lib.rogtk_gpu_last_error(self._ctx).decode())
        return int(p)

    def pinned_free(self, p: int):
        self._lib.rogtk_gpu_pinned_free(self._ctx, ctypes.c_void_p(p))

    def h2d(self, dst: int, src: np.ndarray):
        src = np.ascontiguousarray(src)
        self._check(self._lib.rogtk_gpu_memcpy_h2d(self._ctx, ctypes.c_void_p(dst), src.ctypes.data, src.nbytes))

    def d2h(self, src: int, nbytes: int, dtype=np.uint8) -> np.ndarray:
        out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        self._check(self._lib.rogtk_gpu_memcpy_d2h(self._ctx, out.ctypes.data, ctypes.c_void_p(src), out.nbytes))
        return out

    def synth_reads_device(self, seed: int, umi_base: int, n_umis: int, reads_per_umi: int, read_len: int,
                           template_len: int, err_per_64k: int, read_base: int = 0, umi_ids: np.ndarray | None = None):
        """Generate the SURVEY 8(d) synthetic reads directly in HBM.  Returns (d_data, d_offsets, d_group_offsets,
        n_bytes) as raw device pointers (free with device_free)."""
        n_reads = n_umis * reads_per_umi
        n_bytes = n_reads * read_len
        d_data = self.device_alloc(n_bytes + 16)
        d_off = self.device_alloc((n_reads + 1) * 8)
        d_goff = self.device_alloc((n_umis + 1) * 8)
        d_ids = None
        if umi_ids is not None:   # the UMIs of a hash shard instead of a contiguous range
            ids = np.ascontiguousarray(umi_ids, dtype=np.int64)
            assert len(ids) == n_umis
            d_ids = self.device_alloc(max(ids.nbytes, 16))
            self.h2d(d_ids, ids.view(np.uint8))
        try:
            self._check(self._lib.rogtk_gpu_synth_reads_from(self._ctx, seed & 0xFFFFFFFFFFFFFFFF, umi_base, n_umis, reads_per_umi, read_base,
                                                             read_len, template_len, err_per_64k, ctypes.c_void_p(d_data),
                                                             ctypes.c_void_p(d_off), ctypes.c_void_p(d_goff), ctypes.c_void_p(d_ids) if d_ids else None))
        finally:
            if d_ids:
                self.device_free(d_ids)
        return d_data, d_off, d_goff, n_bytes

    def synth_reads_device_var(self, seed: int, umi_base: int, reads_per_umi: np.ndarray, read_len: int, template_len: int, err_per_64k: int):
        """synth_reads_device with a read count per UMI (BASELINE configs[2]); same bytes as synth.synth_reads."""
        rp = np.asarray(reads_per_umi, dtype=np.int64)
        goff = np.zeros(len(rp) + 1, dtype=np.int64)
        np.cumsum(rp, out=goff[1:])
        n_reads = int(goff[-1])
        n_bytes = n_reads * read_len
        d_data = self.device_alloc(n_bytes + 16)
        d_off = self.device_alloc((n_reads + 1) * 8)
        d_goff = self.device_alloc((len(rp) + 1) * 8)
        self.h2d(d_goff, goff.view(np.uint8))
        self._check(self._lib.rogtk_gpu_synth_reads_var(self._ctx, seed & 0xFFFFFFFFFFFFFFFF, umi_base, len(rp), ctypes.c_void_p(d_goff), n_reads,
                                                        read_len, template_len, err_per_64k, ctypes.c_void_p(d_data), ctypes.c_void_p(d_off)))
        return d_data, d_off, d_goff, n_bytes, n_reads

    # ---- the hot path ----------------------------------------------------------------------------------------------------
    def _call(self, col: _ffi.StrCol, goff_ptr, n_groups: int, k: int, min_coverage: int, flags: int):
        out = _ffi.GraphBatchC()
        rc = self._lib.rogtk_gpu_count_prune_batch(self._ctx, ctypes.byref(col), goff_ptr, n_groups, int(k),
                                                   int(min_coverage), int(flags), ctypes.byref(out))
        self._check(rc)
        return out

    def _to_numpy(self, c: _ffi.GraphBatchC) -> GraphBatch:
        def arr(ptr, n, dt):
            if not ptr or n == 0:
                return np.zeros(0, dtype=dt)
            ct = np.ctypeslib.as_ctypes_type(np.dtype(dt))
            return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ct)), shape=(n,)).copy()

        ng, nk = int(c.n_groups), int(c.n_kmers)
        res = GraphBatch(
            k_eff=int(c.k_eff), n_observations=int(c.n_observations),
            kmer_offsets=arr(c.kmer_offsets, ng + 1, np.int64), kmers_lo=arr(c.kmers_lo, nk, np.uint64),
            kmers_hi=arr(c.kmers_hi, nk, np.uint64) if c.kmers_hi else None,
            exts=arr(c.exts, nk, np.uint8), counts=arr(c.counts, nk, np.uint16),
            stats=arr(c.stats, 3 * ng, np.uint32).reshape(ng, 3) if c.stats else None,
            csr_row_ptr=arr(c.csr_row_ptr, nk + 1, np.int64) if c.csr_row_ptr else None,
            csr_col=arr(c.csr_col, int(c.n_edges), np.int32) if c.csr_row_ptr else None,
            ms_total=float(c.ms_total), ms_count=float(c.ms_count), ms_hot_kernel=float(c.ms_hot_kernel),
            n_spilled_groups=int(c.n_spilled_groups), n_multipass_groups=int(c.n_multipass_groups))
        self._lib.rogtk_gpu_free_batch(self._ctx, ctypes.byref(c))
        return res

    def count_prune(self, data: np.ndarray, offsets: np.ndarray, group_offsets: np.ndarray, k: int, min_coverage: int,
                    validity: np.ndarray | None = None, validity_bit_offset: int = 0, flags: int = 0) -> GraphBatch:
        """Host buffers in, host arrays out (H2D and D2H inside the call).  offsets may be int32 (Utf8) or int64
        (LargeUtf8)."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets)
        if offsets.dtype == np.int32:
            layout = _ffi.UTF8_I32
        else:
            offsets = offsets.astype(np.int64, copy=False)
            layout = _ffi.UTF8_I64
        group_offsets = np.ascontiguousarray(group_offsets, dtype=np.int64)
        if validity is not None:
            validity = np.ascontiguousarray(validity, dtype=np.uint8)
        dptr = (ctypes.c_void_p * 1)(data.ctypes.data if data.size else None)
        dlen = (ctypes.c_int64 * 1)(data.size)
        col = _ffi.StrCol(layout=layout, n_buffers=1, n_rows=len(offsets) - 1,
                          validity=validity.ctypes.data if validity is not None else None,
                          validity_bit_offset=validity_bit_offset, offsets_or_views=offsets.ctypes.data,
                          data=dptr, data_len=dlen)
        c = self._call(col, group_offsets.ctypes.data, len(group_offsets) - 1, k, min_coverage,
                       flags & ~(_ffi.INPUT_ON_DEVICE | _ffi.OUTPUT_ON_DEVICE))
        return self._to_numpy(c)

    def count_prune_views(self, views: np.ndarray, buffers: list[np.ndarray], group_offsets: np.ndarray, k: int,
                          min_coverage: int, validity: np.ndarray | None = None, flags: int = 0) -> GraphBatch:
        """Arrow Utf8View input (Polars >= 0.43 String columns): views = uint8[n_rows,16], buffers = data buffers."""
        views = np.ascontiguousarray(views, dtype=np.uint8).reshape(-1, 16)
        buffers = [np.ascontiguousarray(b, dtype=np.uint8) for b in buffers]
        group_offsets = np.ascontiguousarray(group_offsets, dtype=np.int64)
        nb = len(buffers)
        dptr = (ctypes.c_void_p * max(nb, 1))(*[b.ctypes.data if b.size else None for b in buffers])
        dlen = (ctypes.c_int64 * max(nb, 1))(*[b.size for b in buffers])
        if validity is not None:
            validity = np.ascontiguousarray(validity, dtype=np.uint8)
        col = _ffi.StrCol(layout=_ffi.UTF8_VIEW, n_buffers=nb, n_rows=views.shape[0],
                          validity=validity.ctypes.data if validity is not None else None, validity_bit_offset=0,
                          offsets_or_views=views.ctypes.data, data=dptr, data_len=dlen)
        c = self._call(col, group_offsets.ctypes.data, len(group_offsets) - 1, k, min_coverage,
                       flags & ~(_ffi.INPUT_ON_DEVICE | _ffi.OUTPUT_ON_DEVICE))
        return self._to_numpy(c)

    def count_prune_device(self, d_data: int, n_bytes: int, d_offsets: int, n_rows: int, d_group_offsets: int,
                           n_groups: int, k: int, min_coverage: int, flags: int = 0) -> DeviceBatch:
        """Inputs already resident in HBM (int64 offsets), outputs left in HBM.  Used by bench.py's `value`."""
        dptr = (ctypes.c_void_p * 1)(d_data)
        dlen = (ctypes.c_int64 * 1)(n_bytes)
        col = _ffi.StrCol(layout=_ffi.UTF8_I64, n_buffers=1, n_rows=n_rows, validity=None, validity_bit_offset=0,
                          offsets_or_views=d_offsets, data=dptr, data_len=dlen)
        c = self._call(col, ctypes.c_void_p(d_group_offsets), n_groups, k, min_coverage,
                       flags | _ffi.INPUT_ON_DEVICE | _ffi.OUTPUT_ON_DEVICE)
        return DeviceBatch(k_eff=int(c.k_eff), n_groups=int(c.n_groups), n_kmers=int(c.n_kmers), n_edges=int(c.n_edges),
                           n_observations=int(c.n_observations), ms_total=float(c.ms_total), ms_count=float(c.ms_count),
                           ms_hot_kernel=float(c.ms_hot_kernel), n_spilled_groups=int(c.n_spilled_groups),
                           n_multipass_groups=int(c.n_multipass_groups), _c=c, _engine=self)


    def count_prune_to_device(self, data: np.ndarray, offsets: np.ndarray, group_offsets: np.ndarray, k: int, min_coverage: int,
                              validity: np.ndarray | None = None, flags: int = 0, validity_bit_offset: int = 0) -> DeviceBatch:
        """Host Arrow buffers in, result left in HBM (used by the global-spectrum mode before the all-to-all)."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        group_offsets = np.ascontiguousarray(group_offsets, dtype=np.int64)
        if validity is not None:
            validity = np.ascontiguousarray(validity, dtype=np.uint8)
        dptr = (ctypes.c_void_p * 1)(data.ctypes.data if data.size else None)
        dlen = (ctypes.c_int64 * 1)(data.size)
        col = _ffi.StrCol(layout=_ffi.UTF8_I64, n_buffers=1, n_rows=len(offsets) - 1,
                          validity=validity.ctypes.data if validity is not None else None,
                          validity_bit_offset=int(validity_bit_offset),
                          offsets_or_views=offsets.ctypes.data, data=dptr, data_len=dlen)
        c = self._call(col, group_offsets.ctypes.data, len(group_offsets) - 1, k, min_coverage,
                       (flags | _ffi.OUTPUT_ON_DEVICE) & ~_ffi.INPUT_ON_DEVICE)
        return DeviceBatch(k_eff=int(c.k_eff), n_groups=int(c.n_groups), n_kmers=int(c.n_kmers), n_edges=int(c.n_edges),
                           n_observations=int(c.n_observations), ms_total=float(c.ms_total), ms_count=float(c.ms_count),
                           ms_hot_kernel=float(c.ms_hot_kernel), n_spilled_groups=int(c.n_spilled_groups),
                           n_multipass_groups=int(c.n_multipass_groups), _c=c, _engine=self)

    # ---- global k-mer spectrum (one group spanning every rank) -------------------------------------------------------------
    def spectrum_split(self, d_lo: int, d_hi: int | None, n: int, k_eff_: int, n_ranks: int) -> np.ndarray:
        """bounds[d] = first index of the ascending device list owned by a rank >= d (owner = first-4-base bucket range)."""
        b = np.zeros(n_ranks + 1, dtype=np.int64)
        self._check(self._lib.rogtk_gpu_spectrum_split(self._ctx, ctypes.c_void_p(d_lo), ctypes.c_void_p(d_hi or 0), int(n),
                                                       int(k_eff_), int(n_ranks), b.ctypes.data))
        return b

    def spectrum_merge(self, d_lo: int, d_hi: int | None, d_counts: int, d_exts: int, n: int, k_eff_: int, min_coverage: int,
                       flags: int = 0, on_device: bool = False):
        """Merge (k-mer, partial count, exts) triples received from all ranks: one ascending group with count >= min_coverage.
        on_device=True leaves the result in HBM (DeviceBatch): the carry between the rounds of a chunked global spectrum."""
        out = _ffi.GraphBatchC()
        f = (int(flags) | _ffi.OUTPUT_ON_DEVICE) if on_device else (int(flags) & ~_ffi.OUTPUT_ON_DEVICE)
        self._check(self._lib.rogtk_gpu_spectrum_merge(self._ctx, ctypes.c_void_p(d_lo), ctypes.c_void_p(d_hi or 0),
                                                       ctypes.c_void_p(d_counts), ctypes.c_void_p(d_exts), int(n), int(k_eff_),
                                                       int(min_coverage), f, ctypes.byref(out)))
        if not on_device:
            return self._to_numpy(out)
        return DeviceBatch(k_eff=int(out.k_eff), n_groups=int(out.n_groups), n_kmers=int(out.n_kmers), n_edges=int(out.n_edges),
                           n_observations=int(out.n_observations), ms_total=float(out.ms_total), ms_count=float(out.ms_count),
                           ms_hot_kernel=float(out.ms_hot_kernel), n_spilled_groups=0, n_multipass_groups=0, _c=out, _engine=self)

    # ---- threshold re-use -------------------------------------------------------------------------------------------
    def rethreshold(self, batch: DeviceBatch, min_coverage: int, flags: int = 0) -> GraphBatch:
        """The batch the hot path would return for a higher min_coverage, from a device-resident batch counted once at a lower
        one (sweep_assembly_params / optimize_assembly re-count all reads per cell in the reference)."""
        out = _ffi.GraphBatchC()
        self._check(self._lib.rogtk_gpu_rethreshold(self._ctx, ctypes.byref(batch._c), int(min_coverage),
                                                    int(flags) & ~_ffi.OUTPUT_ON_DEVICE, ctypes.byref(out)))
        return self._to_numpy(out)

    # ---- unitig compaction + contig selection ------------------------------------------------------------------------
    def compress_unitigs(self, batch: DeviceBatch, min_length: int = 0, sizes_only: bool = False, node_parallel: bool = False):
        """Unitigs (maximal non-branching paths, saturating-summed counts) of a device-resident batch and, per group, the
        contig assemble_sequences_expr(method="compression") returns (longest with at least min_length bases).
        sizes_only=True leaves the result in HBM, frees it and returns (n_unitigs, n_bases, device ms): timing without D2H.
        node_parallel=True forces the node-parallel kernels (used when a group exceeds one CTA's shared memory) for testing."""
        out = _ffi.ContigBatchC()
        f = (_ffi.OUTPUT_ON_DEVICE if sizes_only else 0) | (_ffi.UNITIG_NODE_PARALLEL if node_parallel else 0)
        self._check(self._lib.rogtk_gpu_compress_unitigs(self._ctx, ctypes.byref(batch._c), int(min_length), f, ctypes.byref(out)))
        if sizes_only:
            r = (int(out.n_unitigs), int(out.n_bases), float(out.ms_total))
            self._lib.rogtk_gpu_free_contigs(self._ctx, ctypes.byref(out))
            return r

        def arr(ptr, n, dt):
            if n == 0 or not ptr:
                return np.zeros(0, dtype=dt)
            ct = np.ctypeslib.as_ctypes_type(np.dtype(dt))
            return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ct)), shape=(n,)).copy()

        ng, nu, nb = int(out.n_groups), int(out.n_unitigs), int(out.n_bases)
        res = ContigBatch(k_eff=int(out.k_eff), unitig_offsets=arr(out.unitig_offsets, ng + 1, np.int64),
                          seq_offsets=arr(out.seq_offsets, nu + 1, np.int64), seq=arr(out.seq, nb, np.uint8),
                          counts=arr(out.counts, nu, np.uint16), exts=arr(out.exts, nu, np.uint8), cyclic=arr(out.cyclic, nu, np.uint8),
                          largest=arr(out.largest, ng, np.int64), largest_flags=arr(out.largest_flags, ng, np.uint8),
                          ms_total=float(out.ms_total))
        self._lib.rogtk_gpu_free_contigs(self._ctx, ctypes.byref(out))
        return res

    def assemble(self, data: np.ndarray, offsets: np.ndarray, group_offsets: np.ndarray, k: int, min_coverage: int,
                 min_length: int = 0, validity: np.ndarray | None = None, node_parallel: bool = False) -> ContigBatch:
        """reads -> pruned de Bruijn nodes -> unitigs -> the plugin's contig per group, all on the device
        (assemble_sequences_expr with method="compression", reference src/expressions.rs:695-762)."""
        db = self.count_prune_to_device(data, offsets, group_offsets, k, min_coverage, validity=validity)
        try:
            if db.k_eff < 0:   # k > 64: no contigs (reference src/fracture.rs:212-215)
                ng = len(group_offsets) - 1
                z = np.zeros
                return ContigBatch(k_eff=-1, unitig_offsets=z(ng + 1, np.int64), seq_offsets=z(1, np.int64), seq=z(0, np.uint8),
                                   counts=z(0, np.uint16), exts=z(0, np.uint8), cyclic=z(0, np.uint8),
                                   largest=np.full(ng, -1, np.int64), largest_flags=z(ng, np.uint8))
            return self.compress_unitigs(db, min_length, node_parallel=node_parallel)
        finally:
            db.free()

    # ---- whole-expression entries: the reference's Python API (rogtk/__init__.py:104-324), one call per frame -----------------------------
    def _host_col(self, data, offsets, validity=None, validity_bit_offset=0):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets)
        if offsets.dtype == np.int32:
            layout = _ffi.UTF8_I32
        else:
            offsets = offsets.astype(np.int64, copy=False)
            layout = _ffi.UTF8_I64
        if validity is not None:
            validity = np.ascontiguousarray(validity, dtype=np.uint8)
        dptr = (ctypes.c_void_p * 1)(data.ctypes.data if data.size else None)
        dlen = (ctypes.c_int64 * 1)(data.size)
        col = _ffi.StrCol(layout=layout, n_buffers=1, n_rows=len(offsets) - 1,
                          validity=validity.ctypes.data if validity is not None else None,
                          validity_bit_offset=int(validity_bit_offset), offsets_or_views=offsets.ctypes.data, data=dptr, data_len=dlen)
        return col, (data, offsets, validity, dptr, dlen)

    @staticmethod
    def _strings_out(c: "_ffi.StringBatchC") -> "StringBatch":
        n, nb = int(c.n_rows), int(c.n_bytes)

        def arr(ptr, cnt, dt):
            if not ptr or cnt == 0:
                return np.zeros(0, dtype=dt)
            ct = np.ctypeslib.as_ctypes_type(np.dtype(dt))
            return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ct)), shape=(cnt,)).copy()

        off = arr(c.offsets, n + 1, np.int64) if n else np.zeros(1, np.int64)
        return StringBatch(offsets=off, data=arr(c.data, nb, np.uint8), flags=arr(c.flags, n, np.uint8), ms_device=float(c.ms_device),
                           ms_host=float(c.ms_host), n_count_passes=int(c.n_count_passes))

    def assemble_sequences(self, data, offsets, group_offsets, k: int = 10, min_coverage: int = 5, method: str = "shortest_path",
                           start_anchor: str | None = None, end_anchor: str | None = None, min_length: int | None = None,
                           only_largest: bool = True, auto_k: bool = False, validity=None, validity_bit_offset: int = 0,
                           anchors_per_group=None, node_rank=None, host_threads: int = 0, flags: int = 0) -> "StringBatch":
        """rogtk.assemble_sequences / assemble_sequences_with_anchors (reference rogtk/__init__.py:104-234 -> src/expressions.rs:695-849)
        for every group of a frame: one row per group with the contig ("" where the reference returns none).  The expression
        hard-wires only_largest=True (src/expressions.rs:748); pass only_largest=False for fracture_sequences' listing
        (src/fracture.rs:506-564).  anchors_per_group = (start list, end list) gives every group its own anchors."""
        col, keep = self._host_col(data, offsets, validity, validity_bit_offset)
        goff = np.ascontiguousarray(group_offsets, dtype=np.int64)
        if method not in _ffi.METHODS:
            raise RogtkGpuError(-1, "Invalid assembly method. Must be 'compression', 'shortest_path', or 'shortest_path_auto'")
        p = _ffi.AssemblyParamsC(k=int(k), min_coverage=int(min_coverage), method=_ffi.METHODS[method],
                                 start_anchor=start_anchor.encode() if start_anchor is not None else None,
                                 end_anchor=end_anchor.encode() if end_anchor is not None else None,
                                 min_length=int(min_length or 0), only_largest=int(bool(only_largest)), auto_k=int(bool(auto_k)),
                                 host_threads=int(host_threads))
        if anchors_per_group is not None:
            sl, el = anchors_per_group
            so = np.concatenate([[0], np.cumsum([len(x) for x in sl])]).astype(np.int64)
            eo = np.concatenate([[0], np.cumsum([len(x) for x in el])]).astype(np.int64)
            sd = np.frombuffer("".join(sl).encode() or b"\0", dtype=np.uint8).copy()
            ed = np.frombuffer("".join(el).encode() or b"\0", dtype=np.uint8).copy()
            p.start_anchor_offsets, p.start_anchor_data = so.ctypes.data, sd.ctypes.data
            p.end_anchor_offsets, p.end_anchor_data = eo.ctypes.data, ed.ctypes.data
            keep = keep + (so, eo, sd, ed)
        if node_rank is not None:
            node_rank = np.ascontiguousarray(node_rank, dtype=np.int32)
            p.node_rank = node_rank.ctypes.data
        out = _ffi.StringBatchC()
        self._check(self._lib.rogtk_gpu_assemble_batch(self._ctx, ctypes.byref(col), goff.ctypes.data, len(goff) - 1, ctypes.byref(p),
                                                       int(flags) & ~(_ffi.INPUT_ON_DEVICE | _ffi.OUTPUT_ON_DEVICE), ctypes.byref(out)))
        res = self._strings_out(out)
        self._lib.rogtk_gpu_free_strings(self._ctx, ctypes.byref(out))
        return res

    def sweep_assembly_params(self, data, offsets, group_offsets, k_start: int = 5, k_end: int = 64, k_step: int = 1, cov_start: int = 1,
                              cov_end: int = 150, cov_step: int = 1, method: str = "shortest_path", start_anchor: str | None = None,
                              end_anchor: str | None = None, validity=None, validity_bit_offset: int = 0, host_threads: int = 0):
        """rogtk.sweep_assembly_params (reference rogtk/__init__.py:236-287 -> src/expressions.rs:880-955) for every group: returns
        dict(k, min_coverage, contig_length: int64 [n_groups, n_cells], n_count_passes, ms_device, ms_host)."""
        col, keep = self._host_col(data, offsets, validity, validity_bit_offset)
        goff = np.ascontiguousarray(group_offsets, dtype=np.int64)
        if method not in _ffi.METHODS:
            raise RogtkGpuError(-1, f"Invalid assembly method: Unknown assembly method: {method}")
        p = _ffi.SweepParamsC(k_start=k_start, k_end=k_end, k_step=k_step, cov_start=cov_start, cov_end=cov_end, cov_step=cov_step,
                              method=_ffi.METHODS[method], host_threads=int(host_threads),
                              start_anchor=start_anchor.encode() if start_anchor is not None else None,
                              end_anchor=end_anchor.encode() if end_anchor is not None else None)
        out = _ffi.SweepResultC()
        self._check(self._lib.rogtk_gpu_sweep_batch(self._ctx, ctypes.byref(col), goff.ctypes.data, len(goff) - 1, ctypes.byref(p), 0,
                                                    ctypes.byref(out)))
        ng, nc = int(out.n_groups), int(out.n_cells)

        def arr(ptr):
            if ng * nc == 0:
                return np.zeros((ng, nc), np.int64)
            return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_int64)), shape=(ng * nc,)).copy().reshape(ng, nc)

        res = dict(k=arr(out.k), min_coverage=arr(out.min_coverage), contig_length=arr(out.contig_length),
                   n_count_passes=int(out.n_count_passes), ms_device=float(out.ms_device), ms_host=float(out.ms_host))
        self._lib.rogtk_gpu_free_sweep(self._ctx, ctypes.byref(out))
        return res

    def optimize_assembly(self, data, offsets, group_offsets, start_anchor: str, end_anchor: str, method: str = "shortest_path",
                          start_k: int = 31, start_min_coverage: int = 1, max_iterations: int | None = None, explore_k: bool = False,
                          prioritize_length: bool = False, validity=None, validity_bit_offset: int = 0, host_threads: int = 0):
        """rogtk.optimize_assembly (reference rogtk/__init__.py:289-324 -> src/fracture_opt.rs:120-367) for every group: returns
        dict(contig: StringBatch, k, min_coverage, length, input_sequences: uint32 [n_groups], n_cells_evaluated, n_probes)."""
        col, keep = self._host_col(data, offsets, validity, validity_bit_offset)
        goff = np.ascontiguousarray(group_offsets, dtype=np.int64)
        if method not in _ffi.METHODS:
            raise RogtkGpuError(-1, f"Unknown assembly method: {method}")
        p = _ffi.OptimizeParamsC(start_k=start_k, start_min_coverage=start_min_coverage, max_iterations=int(max_iterations or 0),
                                 explore_k=int(bool(explore_k)), prioritize_length=int(bool(prioritize_length)), method=_ffi.METHODS[method],
                                 host_threads=int(host_threads),
                                 start_anchor=start_anchor.encode() if start_anchor is not None else None,
                                 end_anchor=end_anchor.encode() if end_anchor is not None else None)
        out = _ffi.OptimizeResultC()
        self._check(self._lib.rogtk_gpu_optimize_batch(self._ctx, ctypes.byref(col), goff.ctypes.data, len(goff) - 1, ctypes.byref(p), 0,
                                                       ctypes.byref(out)))
        ng = len(goff) - 1

        def arr(ptr):
            if ng == 0:
                return np.zeros(0, np.uint32)
            return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint32)), shape=(ng,)).copy()

        res = dict(contig=self._strings_out(out.contig), k=arr(out.k), min_coverage=arr(out.min_coverage), length=arr(out.length),
                   input_sequences=arr(out.input_sequences), n_cells_evaluated=int(out.n_cells_evaluated), n_probes=int(out.n_probes))
        self._lib.rogtk_gpu_free_optimize(self._ctx, ctypes.byref(out))
        return res

    # ---- Arrow-native entry: one call per frame ----------------------------------------------------------------------
    def _arrow_col(self, child):
        """rogtk_strcol over the buffers of a pyarrow String / LargeString / Binary / StringView array -- no re-encoding, no copy
        (a sliced view array is re-based, pyarrow keeps its variadic buffers whole).  Returns (StrCol, keep-alive objects)."""
        import pyarrow as pa
        ct = child.type
        bufs = child.buffers()
        validity = np.frombuffer(bufs[0], dtype=np.uint8) if bufs[0] is not None and child.null_count else None
        if pa.types.is_string(ct) or pa.types.is_large_string(ct) or pa.types.is_binary(ct) or pa.types.is_large_binary(ct):
            sdt = np.int64 if (pa.types.is_large_string(ct) or pa.types.is_large_binary(ct)) else np.int32
            offs = np.frombuffer(bufs[1], dtype=sdt)[child.offset:child.offset + len(child) + 1] if bufs[1] is not None else np.zeros(1, sdt)
            data = np.frombuffer(bufs[2], dtype=np.uint8) if bufs[2] is not None else np.zeros(0, np.uint8)
            col, keep = self._host_col(data, offs, validity, child.offset)
            return col, keep + (child,)
        if hasattr(pa.types, "is_string_view") and (pa.types.is_string_view(ct) or pa.types.is_binary_view(ct)):
            if child.offset:
                child = pa.concat_arrays([child])   # re-base a sliced view array
                bufs = child.buffers()
                validity = np.frombuffer(bufs[0], dtype=np.uint8) if bufs[0] is not None and child.null_count else None
            views = np.frombuffer(bufs[1], dtype=np.uint8)[:len(child) * 16].reshape(-1, 16) if bufs[1] is not None else np.zeros((0, 16), np.uint8)
            # (a null variadic buffer keeps its place: the views name buffers by index)
            data_bufs = [np.frombuffer(b_, dtype=np.uint8) if b_ is not None else np.zeros(0, np.uint8) for b_ in bufs[2:]]
            nb = len(data_bufs)
            dptr = (ctypes.c_void_p * max(nb, 1))(*[b_.ctypes.data if b_.size else None for b_ in data_bufs])
            dlen = (ctypes.c_int64 * max(nb, 1))(*[b_.size for b_ in data_bufs])
            col = _ffi.StrCol(layout=_ffi.UTF8_VIEW, n_buffers=nb, n_rows=views.shape[0],
                              validity=validity.ctypes.data if validity is not None else None, validity_bit_offset=0,
                              offsets_or_views=views.ctypes.data, data=dptr, data_len=dlen)
            return col, (views, data_bufs, validity, dptr, dlen, child)
        raise TypeError(f"unsupported child type {ct}")

    def assemble_arrow(self, reads, k: int, min_coverage: int, min_length: int = 0, method: str = "compression",
                       start_anchor: str | None = None, end_anchor: str | None = None, auto_k: bool = False):
        """The batched form of assemble_sequences_expr (reference src/expressions.rs:695-762): `reads` is a pyarrow List(String)
        array with one row per group (String / LargeString / StringView children go to the C ABI as they are); returns (pyarrow
        LargeStringArray with the contig of every group, "" where the reference returns no contig; uint8 flags per group: bit 0
        tie in length, bit 1 cyclic -- recompute those groups with the reference if its tie-break matters).  One call of
        rogtk_gpu_assemble_batch: reads -> nodes -> unitigs / paths -> choice; the result buffers become the Arrow array as they
        are."""
        import pyarrow as pa
        goff, child = self._arrow_groups(reads)
        col, keep = self._arrow_col(child)
        if method not in _ffi.METHODS:
            raise RogtkGpuError(-1, "Invalid assembly method. Must be 'compression', 'shortest_path', or 'shortest_path_auto'")
        p = _ffi.AssemblyParamsC(k=int(k), min_coverage=int(min_coverage), method=_ffi.METHODS[method],
                                 start_anchor=start_anchor.encode() if start_anchor is not None else None,
                                 end_anchor=end_anchor.encode() if end_anchor is not None else None,
                                 min_length=int(min_length or 0), only_largest=1, auto_k=int(bool(auto_k)))
        out = _ffi.StringBatchC()
        self._check(self._lib.rogtk_gpu_assemble_batch(self._ctx, ctypes.byref(col), goff.ctypes.data, len(goff) - 1, ctypes.byref(p), 0,
                                                       ctypes.byref(out)))
        res = self._strings_out(out)
        self._lib.rogtk_gpu_free_strings(self._ctx, ctypes.byref(out))
        del keep
        arr = pa.LargeStringArray.from_buffers(len(goff) - 1, pa.py_buffer(res.offsets), pa.py_buffer(res.data))
        return arr, res.flags

    def _arrow_groups(self, reads):
        """(group_offsets int64, child string array) of a List(String) array; null lists become empty groups."""
        import pyarrow as pa
        if isinstance(reads, pa.ChunkedArray):
            reads = reads.combine_chunks() if reads.num_chunks != 1 else reads.chunk(0)
        if not (pa.types.is_list(reads.type) or pa.types.is_large_list(reads.type)):
            raise TypeError("expected a List(String) / LargeList(String) array (one row per group)")
        n_groups = len(reads)
        odt = np.int64 if pa.types.is_large_list(reads.type) else np.int32
        lbufs = reads.buffers()
        goff = np.frombuffer(lbufs[1], dtype=odt)[reads.offset:reads.offset + n_groups + 1].astype(np.int64)
        if reads.null_count:   # a null list is an empty group
            valid = np.unpackbits(np.frombuffer(lbufs[0], dtype=np.uint8), bitorder="little")[reads.offset:reads.offset + n_groups]
            lens = np.diff(goff) * valid
            keep_rows = np.repeat(valid.astype(bool), np.diff(goff))
            vals = reads.values.slice(int(goff[0]), int(goff[-1] - goff[0]))
            if hasattr(pa.types, "is_string_view") and (pa.types.is_string_view(vals.type) or pa.types.is_binary_view(vals.type)):
                vals = vals.cast(pa.large_string())   # pyarrow has no filter kernel for view arrays
            child = vals.filter(pa.array(keep_rows))
            goff = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        else:
            child = reads.values
        return goff, child

    def count_prune_arrow(self, reads, k: int, min_coverage: int, flags: int = 0) -> GraphBatch:
        """`reads` is a pyarrow List/LargeList of String/LargeString/StringView (one row per UMI group: what
        `df.group_by("umi").agg(pl.col("seq"))` hands to a batched plugin call), or a ChunkedArray of those.  The list
        offsets are the group offsets; the child string buffers go to the C ABI without re-encoding.  Null lists are empty
        groups, null strings are dropped (reference src/expressions.rs:737-740)."""
        goff, child = self._arrow_groups(reads)
        col, keep = self._arrow_col(child)
        c = self._call(col, goff.ctypes.data, len(goff) - 1, k, min_coverage, flags & ~(_ffi.INPUT_ON_DEVICE | _ffi.OUTPUT_ON_DEVICE))
        del keep
        return self._to_numpy(c)


def count_prune_multi(engines: list, data, offsets, group_offsets, k: int, min_coverage: int, umi_ids=None, validity=None,
                      validity_bit_offset: int = 0, flags: int = 0):
    """One host frame on several GPUs of this process (rogtk_gpu_count_prune_multi; SURVEY 8e, BASELINE configs[3]): whole groups go to
    the engine rogtk_gpu_shard_of(umi_id) names, every engine counts its share on its own device, the batch comes back in the
    original group order.  Returns (GraphBatch, shard_of_group int32[n_groups])."""
    e0 = engines[0]
    col, keep = e0._host_col(data, offsets, validity, validity_bit_offset)
    goff = np.ascontiguousarray(group_offsets, dtype=np.int64)
    ng = len(goff) - 1
    ctxs = (ctypes.c_void_p * len(engines))(*[e._ctx for e in engines])
    ids = np.ascontiguousarray(umi_ids, dtype=np.uint64) if umi_ids is not None else None
    shard = np.zeros(max(ng, 1), dtype=np.int32)
    out = _ffi.GraphBatchC()
    rc = e0._lib.rogtk_gpu_count_prune_multi(ctxs, len(engines), ctypes.byref(col), goff.ctypes.data, ng,
                                             ids.ctypes.data if ids is not None else None, int(k), int(min_coverage),
                                             int(flags) & ~(_ffi.INPUT_ON_DEVICE | _ffi.OUTPUT_ON_DEVICE), ctypes.byref(out), shard.ctypes.data)
    e0._check(rc)
    return e0._to_numpy(out), shard[:ng]
