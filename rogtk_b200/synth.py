"""Synthetic reads of SURVEY.md 8(d), numpy side.  Bit-identical to the device generator k_synth
(rogtk_b200/csrc/kernels_post.cuh) so that parity subsets can be produced on the host and full-size inputs in HBM.

Per UMI u (global index):  s_u = SM(seed, u)   where SM(seed, c) = mix(seed + (c+1)*GAMMA) is the c-th SplitMix64 output
  template base i        = (SM(s_u, i>>5) >> 2*(i&31)) & 3
  read r start           = SM(s_u, 2^20 + 4096 r) mod (T-L+1)
  read r base j error    : v = 16-bit field (j&3) of SM(s_u, 2^20 + 4096 r + 1 + (j>>2)); if v < err: base = (base+1+v%3)&3
"""
from __future__ import annotations

import numpy as np

GAMMA = np.uint64(0x9E3779B97F4A7C15)
SEED0 = 0x726F67746B  # "rogtk"


def config_seed(config_id: int) -> int:
    return SEED0 ^ config_id


def _mix(z: np.ndarray) -> np.ndarray:
    z = z.astype(np.uint64, copy=True)
    z ^= z >> np.uint64(30)
    z *= np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(27)
    z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    return z


def _sm(seed, c) -> np.ndarray:
    with np.errstate(over="ignore"):
        return _mix(np.asarray(seed, dtype=np.uint64) + (np.asarray(c, dtype=np.uint64) + np.uint64(1)) * GAMMA)


def synth_reads(seed: int, umi_base: int, n_umis: int, reads_per_umi, read_len: int, template_len: int,
                err_per_64k: int = 655):
    """Returns (data uint8, offsets int64, group_offsets int64).  reads_per_umi: int or int array [n_umis]."""
    L, T = int(read_len), int(template_len)
    rp = np.broadcast_to(np.asarray(reads_per_umi, dtype=np.int64), (n_umis,))
    goff = np.zeros(n_umis + 1, dtype=np.int64)
    np.cumsum(rp, out=goff[1:])
    n_reads = int(goff[-1])
    umi = np.repeat(np.arange(n_umis, dtype=np.int64), rp)                     # [n_reads]
    r = np.arange(n_reads, dtype=np.int64) - goff[:-1][umi]                    # read index inside the UMI
    with np.errstate(over="ignore"):
        su = _sm(np.uint64(seed & 0xFFFFFFFFFFFFFFFF), (umi + umi_base).astype(np.uint64))   # [n_reads]
        rc = np.uint64(1 << 20) + np.uint64(4096) * r.astype(np.uint64)
        start = (_sm(su, rc) % np.uint64(T - L + 1)).astype(np.int64)
        j = np.arange(L, dtype=np.int64)[None, :]
        i = start[:, None] + j                                                  # template index [n_reads, L]
        tw = _sm(su[:, None], (i >> 5).astype(np.uint64))
        b = ((tw >> (np.uint64(2) * (i & 31).astype(np.uint64))) & np.uint64(3)).astype(np.int64)
        ew = _sm(su[:, None], rc[:, None] + np.uint64(1) + (j >> 2).astype(np.uint64))
        v = ((ew >> (np.uint64(16) * (j & 3).astype(np.uint64))) & np.uint64(0xFFFF)).astype(np.int64)
    b = np.where(v < err_per_64k, (b + 1 + v % 3) & 3, b)
    data = np.frombuffer(b"ACGT", dtype=np.uint8)[b].reshape(-1)
    offsets = np.arange(n_reads + 1, dtype=np.int64) * L
    return data, offsets, goff


def skewed_reads_per_umi(seed: int, n_umis: int, max_reads: int = 5000) -> np.ndarray:
    """cfg3: reads/UMI = floor(exp(U(0, ln max_reads))) clamped to [1, max_reads] (log-uniform skew)."""
    u = (_sm(np.uint64(seed ^ 0x5EED), np.arange(n_umis, dtype=np.uint64)) >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    return np.clip(np.floor(np.exp(u * np.log(max_reads))), 1, max_reads).astype(np.int64)


def synth_reads_chunked(seed: int, umi_base: int, n_umis: int, reads_per_umi: int, read_len: int, template_len: int,
                        err_per_64k: int = 655, chunk: int = 4096):
    """synth_reads for large n_umis with bounded temporary memory (fixed reads_per_umi)."""
    parts = []
    for a in range(0, n_umis, chunk):
        b = min(n_umis, a + chunk)
        parts.append(synth_reads(seed, umi_base + a, b - a, reads_per_umi, read_len, template_len, err_per_64k)[0])
    data = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    n_reads = n_umis * reads_per_umi
    return data, np.arange(n_reads + 1, dtype=np.int64) * read_len, np.arange(n_umis + 1, dtype=np.int64) * reads_per_umi
