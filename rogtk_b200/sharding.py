"""Host-side UMI -> GPU sharding for the per-UMI path: groups are independent, so each rank takes whole groups
chosen by `splitmix64(umi_id) mod n_ranks` and no collective touches the data path.  Pure numpy; used by the multi-GPU
drivers and covered by the gloo test."""
from __future__ import annotations

import numpy as np

_GAMMA = np.uint64(0x9E3779B97F4A7C15)


def shard_ids(umi_ids: np.ndarray, n_shards: int) -> np.ndarray:
    """Vectorised rogtk_gpu_shard_of (include/rogtk_gpu.h): splitmix64 finaliser of (id + gamma) mod n_shards."""
    if n_shards <= 1:
        return np.zeros(len(umi_ids), dtype=np.int32)
    with np.errstate(over="ignore"):
        z = np.asarray(umi_ids).astype(np.uint64) + _GAMMA
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return (z % np.uint64(n_shards)).astype(np.int32)


def take_groups(data: np.ndarray, offsets: np.ndarray, group_offsets: np.ndarray, groups: np.ndarray):
    """Sub-batch holding the listed groups (in the given order) as contiguous Arrow buffers."""
    offsets = np.asarray(offsets, dtype=np.int64)
    group_offsets = np.asarray(group_offsets, dtype=np.int64)
    rows = np.concatenate([np.arange(group_offsets[g], group_offsets[g + 1]) for g in groups]) if len(groups) else np.zeros(0, np.int64)
    lens = offsets[rows + 1] - offsets[rows] if len(rows) else np.zeros(0, np.int64)
    new_off = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(lens, out=new_off[1:])
    new_data = np.concatenate([data[offsets[r]:offsets[r + 1]] for r in rows]) if len(rows) else np.zeros(0, np.uint8)
    new_goff = np.zeros(len(groups) + 1, dtype=np.int64)
    np.cumsum([group_offsets[g + 1] - group_offsets[g] for g in groups], out=new_goff[1:])
    return new_data.astype(np.uint8), new_off, new_goff


def shard_batch(data, offsets, group_offsets, umi_ids, rank: int, world: int):
    """The groups of this rank (ascending original index) and their sub-batch."""
    mine = np.nonzero(shard_ids(umi_ids, world) == rank)[0]
    return mine, take_groups(data, offsets, group_offsets, mine)
