"""world_size-2 gloo test (CPU): the N>1 driver logic of the per-UMI path — hash sharding of groups, per-rank processing of
whole groups with no data-path collective, reassembly by original group index, and the barrier / max-over-ranks / sum
reductions bench.py uses.  The per-rank engine is stood in by the oracle (no GPU here); the GPU engine itself is covered by
tests/test_gpu_parity.py."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import dbg_oracle
    from rogtk_b200 import sharding, synth
    n = 120
    data, offs, goff = synth.synth_reads(synth.config_seed(1), 0, n, 6, 80, 160)
    umi_ids = np.arange(n, dtype=np.uint64) + 1000
    mine, (d, o, g) = sharding.shard_batch(data, offs, goff, umi_ids, rank, world)
    res = dbg_oracle.count_prune_batch(d, o, g, 17, 2, n_threads=1)
    # scatter this rank's groups back to their original indices
    nsurv = torch.zeros(n, dtype=torch.int64)
    nsurv[torch.from_numpy(mine)] = torch.from_numpy(np.diff(res.kmer_offsets))
    owner = torch.zeros(n, dtype=torch.int64)
    owner[torch.from_numpy(mine)] = 1
    dist.barrier()
    dist.all_reduce(nsurv, op=dist.ReduceOp.SUM)      # sizes only: the k-mers themselves never cross ranks
    dist.all_reduce(owner, op=dist.ReduceOp.SUM)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # bench.py's max-over-ranks timing reduction
    obs = torch.tensor([float(res.n_observations)], dtype=torch.float64)
    dist.all_reduce(obs, op=dist.ReduceOp.SUM)
    np.savez(os.path.join(tmpdir, f"r{rank}.npz"), mine=mine, lo=res.kmers_lo, off=res.kmer_offsets, nsurv=nsurv.numpy(),
             owner=owner.numpy(), tmax=t.numpy(), obs=obs.numpy())
    dist.destroy_process_group()


def test_two_rank_sharding_gloo(tmp_path, oracle):
    import torch.multiprocessing as mp
    from rogtk_b200 import sharding, synth
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    n = 120
    data, offs, goff = synth.synth_reads(synth.config_seed(1), 0, n, 6, 80, 160)
    ref = oracle.count_prune_batch(data, offs, goff, 17, 2)
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    # every group owned by exactly one rank, both ranks non-empty, reductions agree on every rank
    for p in parts:
        assert np.all(p["owner"] == 1) and p["tmax"][0] == world and p["obs"][0] == ref.n_observations
        assert np.array_equal(p["nsurv"], np.diff(ref.kmer_offsets))
    assert all(len(p["mine"]) > n // 4 for p in parts)
    assert sorted(np.concatenate([p["mine"] for p in parts]).tolist()) == list(range(n))
    # reassembled k-mers equal the single-process answer, group by group
    for p in parts:
        for j, g in enumerate(p["mine"]):
            a, b = p["off"][j], p["off"][j + 1]
            ra, rb = ref.kmer_offsets[g], ref.kmer_offsets[g + 1]
            assert np.array_equal(p["lo"][a:b], ref.kmers_lo[ra:rb])
    # the numpy sharding function is the C ABI's rogtk_gpu_shard_of
    import rogtk_b200
    ids = np.arange(500, dtype=np.uint64) * 7919
    for w in (2, 4, 8):
        assert np.array_equal(sharding.shard_ids(ids, w), [rogtk_b200.shard_of(int(i), w) for i in ids])


def _spectrum_worker(rank, world, port, tmpdir):
    """Protocol of the global k-mer-spectrum mode (rogtk_b200/spectrum.py) with the oracle standing in for the GPU engine:
    local distinct k-mers (min_cov=1) -> split by owner -> all-to-all -> merge (saturating sum, OR) -> min_coverage."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rogtk_b200
    from oracle import dbg_oracle
    from rogtk_b200 import synth

    def _exchange(send, send_counts, group, device):
        """all-to-all of a 1-D tensor with per-destination counts over gloo (on the GPU the library does this step itself with grouped
        ncclSend / ncclRecv, rogtk_gpu_spectrum_exchange; the protocol around it is what this CPU test covers)"""
        if send.dtype != torch.uint8:   # payloads travel as bytes (no 16-bit integer type on the wire)
            isz, dt = send.element_size(), send.dtype
            recv, rcnt = _exchange(send.view(torch.uint8), [c * isz for c in send_counts], group, device)
            return recv.view(dt), [c // isz for c in rcnt]
        sc = torch.tensor(send_counts, dtype=torch.int64)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc, group=group)
        recv_counts = [int(x) for x in rc]
        recv = torch.empty(sum(recv_counts), dtype=send.dtype)
        dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=list(send_counts), group=group)
        return recv, recv_counts
    data, offs, _ = synth.synth_reads(synth.config_seed(5), 0, 30, 20, 100, 300)
    n_rows, L = len(offs) - 1, 100
    rows = np.arange(rank, n_rows, world)
    d = data.reshape(n_rows, L)[rows].reshape(-1)
    o = np.arange(len(rows) + 1, dtype=np.int64) * L
    k, mc = 21, 2
    loc = dbg_oracle.count_prune_batch(d, o, np.array([0, len(rows)]), k, 1)
    owner = np.array([rogtk_b200.ffi.load().rogtk_gpu_spectrum_owner(int(x), 0, loc.k_eff, world) for x in loc.kmers_lo])
    assert np.all(np.diff(owner) >= 0)
    send_counts = [int((owner == r).sum()) for r in range(world)]
    r_lo, _ = _exchange(torch.from_numpy(loc.kmers_lo.view(np.int64)), send_counts, None, "cpu")
    r_ct, _ = _exchange(torch.from_numpy(loc.counts.view(np.int16)), send_counts, None, "cpu")
    r_ex, _ = _exchange(torch.from_numpy(loc.exts), send_counts, None, "cpu")
    lo, ct, ex = r_lo.numpy().view(np.uint64), r_ct.numpy().view(np.uint16).astype(np.int64), r_ex.numpy()
    order = np.argsort(lo, kind="stable")
    lo, ct, ex = lo[order], ct[order], ex[order]
    ulo, start = np.unique(lo, return_index=True)
    csum = np.minimum(65535, np.add.reduceat(ct, start)) if len(lo) else ct
    eor = np.bitwise_or.reduceat(ex, start) if len(lo) else ex
    keep = csum >= mc
    np.savez(os.path.join(tmpdir, f"s{rank}.npz"), lo=ulo[keep], ct=csum[keep].astype(np.uint16), ex=eor[keep])
    dist.destroy_process_group()


def test_global_spectrum_protocol_gloo(tmp_path, oracle):
    import torch.multiprocessing as mp
    from rogtk_b200 import synth
    world, port = 2, _free_port()
    mp.spawn(_spectrum_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    data, offs, _ = synth.synth_reads(synth.config_seed(5), 0, 30, 20, 100, 300)
    ref = oracle.count_prune_batch(data, offs, np.array([0, len(offs) - 1]), 21, 2)
    parts = [np.load(tmp_path / f"s{r}.npz") for r in range(world)]
    assert all(len(p["lo"]) > 0 for p in parts)
    assert np.array_equal(np.concatenate([p["lo"] for p in parts]), ref.kmers_lo)      # rank order == ascending order
    assert np.array_equal(np.concatenate([p["ct"] for p in parts]), ref.counts)
    assert np.array_equal(np.concatenate([p["ex"] for p in parts]), ref.exts)
