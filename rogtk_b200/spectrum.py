"""Global k-mer spectrum: ONE group spanning the reads held by every rank (BASELINE configs[4]).

The reference has no multi-process form; its closest call is the same expression on an un-grouped column, i.e. one
`filter_kmers` over all reads (reference src/fracture.rs:110-116), which internally walks the 256 first-4-base buckets in
ascending order.  Here every rank
  1. counts its own reads as one group with min_coverage=1 (distinct k-mers, partial u16-saturated counts, exts) on its GPU,
  2. hands the device-resident result to ``rogtk_gpu_spectrum_exchange``: split at the owner boundaries (owner = contiguous
     range of first-4-base buckets), one grouped ncclSend/ncclRecv per peer of a packed record stream over NVLink, merge of what
     arrived (sum partial counts saturating, OR exts, apply min_coverage) -- all inside the C library.
Rank-ordered concatenation of the per-rank outputs is the single-process answer (ascending), bit for bit.

No torch here: the NCCL communicator belongs to the library (dlopen'ed libnccl); the 128-byte unique id and the two scalars the
driver needs from every rank travel over a plain TCP socket through rank 0 (`Rendezvous`).
"""
from __future__ import annotations

import ctypes
import socket
import struct
import time

import numpy as np

from . import _ffi
from .engine import DeviceBatch, Engine, GraphBatch


class Rendezvous:
    """Star-shaped control channel through rank 0 (a few bytes per call; data never goes here)."""

    def __init__(self, rank: int, world: int, addr: str = "127.0.0.1", port: int = 29577, timeout: float = 120.0):
        self.rank, self.world = rank, world
        self.peers = []
        self.sock = None
        if world == 1:
            return
        if rank == 0:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((addr, port))
            srv.listen(world)
            srv.settimeout(timeout)
            got = {}
            while len(got) < world - 1:
                c, _ = srv.accept()
                c.settimeout(timeout)
                r = struct.unpack("<i", self._recv(c, 4))[0]
                got[r] = c
            self.peers = [got[r] for r in range(1, world)]
            srv.close()
        else:
            t0 = time.time()
            while True:
                try:
                    s = socket.create_connection((addr, port), timeout=5.0)
                    break
                except OSError:
                    if time.time() - t0 > timeout:
                        raise
                    time.sleep(0.1)
            s.settimeout(timeout)
            s.sendall(struct.pack("<i", rank))
            self.sock = s

    @staticmethod
    def _recv(s, n):
        buf = b""
        while len(buf) < n:
            part = s.recv(n - len(buf))
            if not part:
                raise ConnectionError("rendezvous peer closed")
            buf += part
        return buf

    def broadcast(self, payload: bytes | None, n: int) -> bytes:
        if self.world == 1:
            return payload
        if self.rank == 0:
            for p in self.peers:
                p.sendall(payload)
            return payload
        return self._recv(self.sock, n)

    def allreduce(self, value: int, op: str = "sum") -> int:
        if self.world == 1:
            return int(value)
        if self.rank == 0:
            vals = [int(value)] + [struct.unpack("<q", self._recv(p, 8))[0] for p in self.peers]
            r = max(vals) if op == "max" else sum(vals)
            for p in self.peers:
                p.sendall(struct.pack("<q", r))
            return r
        self.sock.sendall(struct.pack("<q", int(value)))
        return struct.unpack("<q", self._recv(self.sock, 8))[0]

    def barrier(self):
        self.allreduce(0)

    def close(self):
        for p in self.peers:
            p.close()
        if self.sock is not None:
            self.sock.close()


class ExchangeStats(ctypes.Structure):
    _fields_ = [("n_sent", ctypes.c_int64), ("n_received", ctypes.c_int64), ("bytes_sent", ctypes.c_int64),
                ("bytes_received", ctypes.c_int64), ("ms_exchange", ctypes.c_float), ("n_ranks", ctypes.c_int32)]


def init_comm(engine: Engine, rdv: Rendezvous):
    """Create the library-owned NCCL communicator of this rank (rogtk_gpu_nccl_init); no-op for one rank."""
    if rdv.world == 1:
        return
    lib = engine._lib
    ident = (ctypes.c_uint8 * 128)()
    if rdv.rank == 0:
        engine._check(lib.rogtk_gpu_nccl_unique_id(ident))
    raw = rdv.broadcast(bytes(ident) if rdv.rank == 0 else None, 128)
    ident = (ctypes.c_uint8 * 128).from_buffer_copy(raw)
    engine._check(lib.rogtk_gpu_nccl_init(engine._ctx, ident, rdv.rank, rdv.world))


def finalize_comm(engine: Engine):
    engine._lib.rogtk_gpu_nccl_finalize(engine._ctx)


def _exchange(engine: Engine, local: DeviceBatch, carry: DeviceBatch | None, min_coverage: int, flags: int, on_device: bool):
    out = _ffi.GraphBatchC()
    st = ExchangeStats()
    f = (int(flags) | _ffi.OUTPUT_ON_DEVICE) if on_device else (int(flags) & ~_ffi.OUTPUT_ON_DEVICE)
    engine._check(engine._lib.rogtk_gpu_spectrum_exchange(engine._ctx, ctypes.byref(local._c), ctypes.byref(carry._c) if carry is not None else None,
                                                          int(min_coverage), f, ctypes.byref(out), ctypes.byref(st)))
    if not on_device:
        return engine._to_numpy(out), st
    return DeviceBatch(k_eff=int(out.k_eff), n_groups=int(out.n_groups), n_kmers=int(out.n_kmers), n_edges=int(out.n_edges),
                       n_observations=int(out.n_observations), ms_total=float(out.ms_total), ms_count=float(out.ms_count),
                       ms_hot_kernel=float(out.ms_hot_kernel), n_spilled_groups=0, n_multipass_groups=0, _c=out, _engine=engine), st


def _device_batch(engine: Engine, out: "_ffi.GraphBatchC") -> DeviceBatch:
    return DeviceBatch(k_eff=int(out.k_eff), n_groups=int(out.n_groups), n_kmers=int(out.n_kmers), n_edges=int(out.n_edges),
                       n_observations=int(out.n_observations), ms_total=float(out.ms_total), ms_count=float(out.ms_count),
                       ms_hot_kernel=float(out.ms_hot_kernel), n_spilled_groups=0, n_multipass_groups=0, _c=out, _engine=engine)


def _merge_levels(engine: Engine, parts: list, min_coverage: int, flags: int, on_device: bool):
    """rogtk_gpu_spectrum_merge_runs over device-resident shares (no communication)."""
    arr = (ctypes.c_void_p * len(parts))(*[ctypes.addressof(p._c) for p in parts])
    out = _ffi.GraphBatchC()
    f = (int(flags) | _ffi.OUTPUT_ON_DEVICE) if on_device else (int(flags) & ~_ffi.OUTPUT_ON_DEVICE)
    engine._check(engine._lib.rogtk_gpu_spectrum_merge_runs(engine._ctx, arr, len(parts), int(min_coverage), f, ctypes.byref(out)))
    return _device_batch(engine, out) if on_device else engine._to_numpy(out)


def _run_rounds(engine: Engine, rdv: Rendezvous, rounds: int, count_round, min_coverage: int, flags: int, keep_on_device: bool,
                stats_out: dict | None):
    """The rounds of the mode.  count_round(r) -> this rank's reads of round r counted as one group (DeviceBatch, min_coverage 1).
    The share of the spectrum this rank owns is carried in HBM as LEVELS.  The exchange of a round merges what arrives with the
    newest level while that level is small (up to 3 rounds' worth: the merge runs inside the exchange call, no extra pass), else
    it leaves a new level; two levels are merged when the newer one has reached 3/4 of the older one.  A carried record is
    re-read O(log rounds) times instead of once per round (20 rounds: about half the merge traffic, 100 rounds: a sixth);
    min_coverage is applied by the last merge."""
    levels: list[DeviceBatch] = []
    n_obs = 0
    unit = None   # records one round adds to this rank's share (first round's result)
    tot = dict(bytes_sent=0, bytes_received=0, ms_exchange=0.0, n_sent=0, n_received=0, rounds=rounds, ms_local=0.0, ms_merge=0.0,
               level_merges=0)
    try:
        for r in range(rounds):
            local = count_round(r)
            try:
                if local.k_eff < 0:   # k > 64: empty result (reference src/fracture.rs:212-215)
                    return local.to_host()
                n_obs += local.n_observations
                tot["ms_local"] += local.ms_total
                last = r == rounds - 1
                carry = levels[-1] if levels and unit is not None and levels[-1].n_kmers <= 3 * unit else None
                closing = last and len(levels) == (1 if carry is not None else 0)    # this exchange produces the final result
                merged, st = _exchange(engine, local, carry, min_coverage if closing else 1, (flags & _ffi.EMIT_STATS) if closing else 0,
                                       on_device=(not closing) or keep_on_device)
                for key in ("bytes_sent", "bytes_received", "n_sent", "n_received"):
                    tot[key] += int(getattr(st, key))
                tot["ms_exchange"] += float(st.ms_exchange)
                tot["ms_merge"] += float(merged.ms_total)
            finally:
                local.free()
            if carry is not None:
                levels.pop().free()
            if closing:
                merged.n_observations = rdv.allreduce(n_obs)   # observations of the whole job (a size, not data)
                if stats_out is not None:
                    stats_out.update(tot)
                return merged
            if unit is None:
                unit = max(1, merged.n_kmers)
            levels.append(merged)
            while len(levels) >= 2 and 4 * levels[-1].n_kmers >= 3 * levels[-2].n_kmers:
                m = _merge_levels(engine, levels[-2:], 1, 0, True)
                tot["ms_merge"] += float(m.ms_total)
                tot["level_merges"] += 1
                levels.pop().free()
                levels.pop().free()
                levels.append(m)
        final = _merge_levels(engine, levels, min_coverage, flags & _ffi.EMIT_STATS, keep_on_device)
        tot["ms_merge"] += float(final.ms_total)
        tot["level_merges"] += 1
        final.n_observations = rdv.allreduce(n_obs)
        if stats_out is not None:
            stats_out.update(tot)
        return final
    finally:
        for lv in levels:
            lv.free()


def global_spectrum_device(engine: Engine, d_data: int, n_bytes: int, d_offsets: int, n_rows: int, k: int, min_coverage: int,
                           rdv: Rendezvous | None = None, flags: int = 0, chunk_rows: int | None = None, stats_out: dict | None = None,
                           keep_on_device: bool = False):
    """global_spectrum for reads already resident in HBM (one LargeUtf8 column: d_data bytes, d_offsets int64[n_rows + 1]).
    Rounds of chunk_rows reads address the same column through an advanced offsets pointer -- nothing is copied.
    keep_on_device=True returns the rank's share as a DeviceBatch (free it), else a host GraphBatch."""
    rdv = rdv or Rendezvous(0, 1)
    step = n_rows if not chunk_rows or chunk_rows <= 0 else int(chunk_rows)
    rounds = rdv.allreduce(max(1, -(-n_rows // max(step, 1))), "max")
    d_goff = engine.device_alloc(16)

    def count_round(r):
        a, b = min(n_rows, r * step), min(n_rows, (r + 1) * step)
        engine.h2d(d_goff, np.array([0, b - a], dtype=np.int64).view(np.uint8))
        return engine.count_prune_device(d_data, n_bytes, d_offsets + 8 * a, b - a, d_goff, 1, k, 1, flags=flags & _ffi.EXACT_K)

    try:
        return _run_rounds(engine, rdv, rounds, count_round, min_coverage, flags, keep_on_device, stats_out)
    finally:
        engine.device_free(d_goff)


def global_spectrum(engine: Engine, data: np.ndarray, offsets: np.ndarray, k: int, min_coverage: int, rdv: Rendezvous | None = None,
                    validity: np.ndarray | None = None, flags: int = 0, chunk_rows: int | None = None, stats_out: dict | None = None) -> GraphBatch:
    """This rank's share of the global spectrum of all ranks' reads.  `rdv` is the control channel of the job (None = one rank);
    call init_comm(engine, rdv) once before.

    `chunk_rows` bounds the reads counted at a time: the local table of a round is sized by the round's reads, not by the
    whole input, and the rank's share of the spectrum (distinct k-mers it owns, counts saturating at 65535, exts) is carried in
    HBM from round to round (in levels, see _run_rounds); min_coverage is applied at the end.  Every rank runs the same number of
    rounds.  stats_out (optional dict) receives the NCCL byte / time totals of the exchanges."""
    rdv = rdv or Rendezvous(0, 1)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n_rows = len(offsets) - 1
    step = n_rows if not chunk_rows or chunk_rows <= 0 else int(chunk_rows)
    rounds = rdv.allreduce(max(1, -(-n_rows // max(step, 1))), "max")
    if validity is not None and rounds > 1:
        raise ValueError("chunk_rows with a validity bitmap is not supported: drop the null rows first")

    def count_round(r):
        a, b = min(n_rows, r * step), min(n_rows, (r + 1) * step)
        d = data[offsets[a]:offsets[b]]
        o = offsets[a:b + 1] - offsets[a]
        return engine.count_prune_to_device(d, o, np.array([0, b - a], dtype=np.int64), k, 1,
                                            validity=validity if rounds == 1 else None, flags=flags & _ffi.EXACT_K)

    return _run_rounds(engine, rdv, rounds, count_round, min_coverage, flags, False, stats_out)
