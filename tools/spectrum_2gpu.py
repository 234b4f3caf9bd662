"""2-GPU (torchrun, NCCL) check + timing of the global k-mer-spectrum mode against the oracle (run on the GPU box):
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/spectrum_2gpu.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import rogtk_b200
from rogtk_b200 import spectrum, synth

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
eng = rogtk_b200.Engine(device=lr)
n_tmpl, reads_per, L, T = int(sys.argv[1]) if len(sys.argv) > 1 else 2000, 50, 150, 500
CHUNK = int(sys.argv[2]) if len(sys.argv) > 2 else None      # reads per round on every rank (None: one round)
data, offs, _ = synth.synth_reads(synth.config_seed(5), 0, n_tmpl, reads_per, L, T)
n_rows = len(offs) - 1
rows = np.arange(rank, n_rows, world)                       # reads striped round-robin over ranks
d = data.reshape(n_rows, L)[rows].reshape(-1)
o = np.arange(len(rows) + 1, dtype=np.int64) * L
out = {}
for k, mc in ((31, 2),):
    for it in range(3):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        got = spectrum.global_spectrum(eng, d, o, k, mc, chunk_rows=CHUNK)
        torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    sizes = [None] * world
    dist.all_gather_object(sizes, (len(got.kmers_lo), got.kmers_lo, got.exts, got.counts))
    if rank == 0:
        from oracle import dbg_oracle
        ref = dbg_oracle.count_prune_batch(data, offs, np.array([0, n_rows]), k, mc)
        lo = np.concatenate([s[1] for s in sizes]); ex = np.concatenate([s[2] for s in sizes]); ct = np.concatenate([s[3] for s in sizes])
        ok = np.array_equal(lo, ref.kmers_lo) and np.array_equal(ex, ref.exts) and np.array_equal(ct, ref.counts)
        print(json.dumps({"mode": "global_spectrum", "world": world, "reads": n_rows, "chunk_rows": CHUNK, "k": k, "min_coverage": mc, "observations": ref.n_observations,
                          "survivors": int(len(ref.kmers_lo)), "per_rank_survivors": [s[0] for s in sizes], "parity_vs_oracle": bool(ok),
                          "wall_ms": dt * 1e3, "obs_per_s": ref.n_observations / dt}))
        assert ok
dist.destroy_process_group()
