"""Global k-mer spectrum: ONE group spanning the reads held by every rank (BASELINE configs[4]).

The reference has no multi-process form; its closest call is the same expression on an un-grouped column, i.e. one
`filter_kmers` over all reads (reference src/fracture.rs:110-116), which internally walks the 256 first-4-base buckets in
ascending order.  Here every rank
  1. counts its own reads as one group with min_coverage=1 (distinct k-mers, partial u16-saturated counts, exts) on its GPU,
  2. splits the ascending result at the owner boundaries (owner = contiguous range of first-4-base buckets),
  3. exchanges the four arrays with one all-to-all (NCCL over NVLink when the process group is NCCL),
  4. merges what it received: sums partial counts (saturating), ORs exts, applies min_coverage.
Rank-ordered concatenation of the per-rank outputs is the single-process answer (ascending), bit for bit.
"""
from __future__ import annotations

import numpy as np

from . import _ffi
from .engine import Engine, GraphBatch


class _DevArray:
    """Zero-copy torch view of a raw device pointer owned by the C library."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def _as_tensor(ptr, n, typestr, torch_dtype, device):
    import torch
    if n == 0:
        return torch.empty(0, dtype=torch_dtype, device=device)
    return torch.as_tensor(_DevArray(ptr, n, typestr), device=device).view(torch_dtype)


def _exchange(send, send_counts, group, device):
    """all-to-all of a 1-D device tensor with per-destination counts; returns (recv tensor, recv_counts).
    Payloads travel as bytes (NCCL has no 16-bit integer type)."""
    import torch
    import torch.distributed as dist
    if send.dtype != torch.uint8:
        isz = send.element_size()
        dt = send.dtype
        recv, rc = _exchange(send.view(torch.uint8), [c * isz for c in send_counts], group, device)
        return recv.view(dt), [c // isz for c in rc]
    world = dist.get_world_size(group)
    sc = torch.tensor(send_counts, dtype=torch.int64, device=device if dist.get_backend(group) == "nccl" else "cpu")
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(x) for x in rc.cpu()]
    if dist.get_backend(group) == "nccl":
        recv = torch.empty(sum(recv_counts), dtype=send.dtype, device=device)
        dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=list(send_counts), group=group)
        return recv, recv_counts
    # gloo (CPU tests / single-GPU emulation): stage through host memory
    s_cpu = send.cpu()
    r_cpu = torch.empty(sum(recv_counts), dtype=send.dtype)
    dist.all_to_all_single(r_cpu, s_cpu, output_split_sizes=recv_counts, input_split_sizes=list(send_counts), group=group)
    return r_cpu.to(device), recv_counts


def global_spectrum(engine: Engine, data: np.ndarray, offsets: np.ndarray, k: int, min_coverage: int, group=None,
                    validity: np.ndarray | None = None, flags: int = 0) -> GraphBatch:
    """This rank's share of the global spectrum of all ranks' reads.  `group` is a torch.distributed process group
    (None = default group; if torch.distributed is not initialised the call is the single-rank form)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    device = torch.device("cuda", engine.device)
    n_rows = len(offsets) - 1
    local = engine.count_prune_to_device(data, offsets, np.array([0, n_rows], dtype=np.int64), k, 1, validity=validity,
                                         flags=flags & _ffi.EXACT_K)
    try:
        K = local.k_eff
        if K < 0:   # k > 64: empty result (reference src/fracture.rs:212-215)
            return local.to_host()
        c, n = local._c, local.n_kmers
        bounds = engine.spectrum_split(c.kmers_lo, c.kmers_hi, n, K, world)
        send_counts = [int(bounds[d + 1] - bounds[d]) for d in range(world)]
        lo = _as_tensor(c.kmers_lo, n, "<i8", torch.int64, device)
        hi = _as_tensor(c.kmers_hi, n, "<i8", torch.int64, device) if K > 32 else None
        cnt = _as_tensor(c.counts, n, "<i2", torch.int16, device)
        ext = _as_tensor(c.exts, n, "|u1", torch.uint8, device)
        if world > 1:
            r_lo, _ = _exchange(lo, send_counts, group, device)
            r_hi = _exchange(hi, send_counts, group, device)[0] if hi is not None else None
            r_cnt, _ = _exchange(cnt, send_counts, group, device)
            r_ext, _ = _exchange(ext, send_counts, group, device)
        else:
            r_lo, r_hi, r_cnt, r_ext = lo, hi, cnt, ext
        torch.cuda.synchronize(device)
        n_obs = local.n_observations
        if world > 1:   # observations of the whole job (a size, not data)
            t = torch.tensor([n_obs], dtype=torch.int64, device=device if dist.get_backend(group) == "nccl" else "cpu")
            dist.all_reduce(t, group=group)
            n_obs = int(t.item())
        res = engine.spectrum_merge(r_lo.data_ptr() if r_lo.numel() else 0, r_hi.data_ptr() if r_hi is not None and r_hi.numel() else 0,
                                     r_cnt.data_ptr() if r_cnt.numel() else 0, r_ext.data_ptr() if r_ext.numel() else 0,
                                     int(r_lo.numel()), K, min_coverage, flags=flags & _ffi.EMIT_STATS)
        res.n_observations = n_obs
        return res
    finally:
        local.free()
