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
                    validity: np.ndarray | None = None, flags: int = 0, chunk_rows: int | None = None) -> GraphBatch:
    """This rank's share of the global spectrum of all ranks' reads.  `group` is a torch.distributed process group
    (None = default group; if torch.distributed is not initialised the call is the single-rank form).

    `chunk_rows` bounds the reads counted at a time: the local table of a round is sized by the round's reads, not by the
    whole input, and the rank's share of the spectrum (distinct k-mers it owns, counts saturating at 65535, exts) is carried in
    HBM from round to round; min_coverage is applied in the last round only.  Every rank runs the same number of rounds."""
    import torch
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if multi else 1
    nccl = multi and dist.get_backend(group) == "nccl"
    device = torch.device("cuda", engine.device)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n_rows = len(offsets) - 1
    step = n_rows if not chunk_rows or chunk_rows <= 0 else int(chunk_rows)
    rounds = max(1, -(-n_rows // max(step, 1)))
    if world > 1:
        t = torch.tensor([rounds], dtype=torch.int64, device=device if nccl else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        rounds = int(t.item())
    if validity is not None and rounds > 1:
        raise ValueError("chunk_rows with a validity bitmap is not supported: drop the null rows first")

    carry = None        # device-resident share of the spectrum after the previous rounds (C batch, min_coverage 1)
    carry_t = None      # torch views of it
    n_obs = 0
    K = None
    try:
        for r in range(rounds):
            a, b = min(n_rows, r * step), min(n_rows, (r + 1) * step)
            d = data[offsets[a]:offsets[b]]
            o = offsets[a:b + 1] - offsets[a]
            local = engine.count_prune_to_device(d, o, np.array([0, b - a], dtype=np.int64), k, 1,
                                                 validity=validity if rounds == 1 else None, flags=flags & _ffi.EXACT_K)
            try:
                K = local.k_eff
                if K < 0:   # k > 64: empty result (reference src/fracture.rs:212-215)
                    return local.to_host()
                n_obs += local.n_observations
                c, n = local._c, local.n_kmers
                bounds = engine.spectrum_split(c.kmers_lo, c.kmers_hi, n, K, world)
                send_counts = [int(bounds[i + 1] - bounds[i]) for i in range(world)]
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
                if carry_t is not None:   # what this rank owns so far joins the merge as one more source
                    r_lo = torch.cat([carry_t[0], r_lo])
                    r_hi = torch.cat([carry_t[1], r_hi]) if r_hi is not None else None
                    r_cnt = torch.cat([carry_t[2], r_cnt])
                    r_ext = torch.cat([carry_t[3], r_ext])
                torch.cuda.synchronize(device)
                last = r == rounds - 1
                ptr = lambda t_: t_.data_ptr() if t_ is not None and t_.numel() else 0
                merged = engine.spectrum_merge(ptr(r_lo), ptr(r_hi), ptr(r_cnt), ptr(r_ext), int(r_lo.numel()), K,
                                               min_coverage if last else 1, flags=flags & _ffi.EMIT_STATS if last else 0,
                                               on_device=not last)
            finally:
                local.free()
            if carry is not None:
                carry.free()
                carry, carry_t = None, None
            if last:
                if world > 1:   # observations of the whole job (a size, not data)
                    t = torch.tensor([n_obs], dtype=torch.int64, device=device if nccl else "cpu")
                    dist.all_reduce(t, group=group)
                    n_obs = int(t.item())
                merged.n_observations = n_obs
                return merged
            carry = merged
            cc, m = carry._c, carry.n_kmers
            carry_t = (_as_tensor(cc.kmers_lo, m, "<i8", torch.int64, device),
                       _as_tensor(cc.kmers_hi, m, "<i8", torch.int64, device) if K > 32 else None,
                       _as_tensor(cc.counts, m, "<i2", torch.int16, device), _as_tensor(cc.exts, m, "|u1", torch.uint8, device))
    finally:
        if carry is not None:
            carry.free()
