"""N-GPU check + timing of the global k-mer-spectrum mode against the oracle (run on the GPU box; one process per GPU, no torch:
the NCCL communicator belongs to librogtk_gpu.so, the unique id travels over a TCP socket).
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/spectrum_ngpu.py [templates] [chunk_rows] [parity 0/1]
(torchrun is only the launcher: it provides RANK / WORLD_SIZE / LOCAL_RANK / MASTER_PORT.)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rogtk_b200
from rogtk_b200 import spectrum, synth

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
port = int(os.environ.get("MASTER_PORT", 29533)) + 17
eng = rogtk_b200.Engine(device=lr)
rdv = spectrum.Rendezvous(rank, world, "127.0.0.1", port)
spectrum.init_comm(eng, rdv)
n_tmpl, reads_per, L, T = int(sys.argv[1]) if len(sys.argv) > 1 else 2000, 50, 150, 500
CHUNK = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else None      # reads per round on every rank (None: one round)
PARITY = int(sys.argv[3]) if len(sys.argv) > 3 else 1
# every rank generates ITS stripe of the reads: templates are striped over the ranks' read streams (reads of template t: rows t*50 ..)
if PARITY:
    data, offs, _ = synth.synth_reads(synth.config_seed(5), 0, n_tmpl, reads_per, L, T)
    n_rows = len(offs) - 1
    rows = np.arange(rank, n_rows, world)                       # reads striped round-robin over ranks
    d = data.reshape(n_rows, L)[rows].reshape(-1)
else:
    # timing runs: every rank draws ITS reads_per / world reads of EVERY template straight into HBM (same templates on all ranks,
    # different reads), and the whole mode runs device-resident: the reads never exist on the host
    per = max(1, reads_per // world)
    d_data, d_off, d_goff, n_bytes = eng.synth_reads_device(synth.config_seed(5), 0, n_tmpl, per, L, T, 655, read_base=rank * per)
    n_local = n_tmpl * per
    n_rows = n_local * world
if PARITY:
    o = np.arange(len(d) // L + 1, dtype=np.int64) * L
k, mc = 31, 2
st = {}
ITERS = int(sys.argv[4]) if len(sys.argv) > 4 else (3 if PARITY else 4)
for it in range(ITERS):
    rdv.barrier(); t0 = time.perf_counter()
    if PARITY:
        got = spectrum.global_spectrum(eng, d, o, k, mc, rdv=rdv, chunk_rows=CHUNK, stats_out=st)
        n_mine = len(got.kmers_lo)
    else:
        got = spectrum.global_spectrum_device(eng, d_data, n_bytes, d_off, n_local, k, mc, rdv=rdv, chunk_rows=CHUNK, stats_out=st,
                                              keep_on_device=True)
        n_mine = got.n_kmers
        got.free()
    rdv.barrier(); dt = time.perf_counter() - t0
tot_surv = rdv.allreduce(n_mine)
max_bytes = rdv.allreduce(st["bytes_sent"], "max")
max_ms = rdv.allreduce(int(st["ms_exchange"] * 1000), "max") / 1000.0
ok = None
if PARITY:
    # parity: ranks write their pieces, rank 0 concatenates in rank order and compares with one un-grouped oracle call
    np.savez(f"/tmp/spec_{rank}.npz", lo=got.kmers_lo, ex=got.exts, ct=got.counts)
    rdv.barrier()
    if rank == 0:
        from oracle import dbg_oracle
        ref = dbg_oracle.count_prune_batch(data, offs, np.array([0, n_rows]), k, mc)
        parts = [np.load(f"/tmp/spec_{r}.npz") for r in range(world)]
        lo = np.concatenate([p["lo"] for p in parts]); ex = np.concatenate([p["ex"] for p in parts]); ct = np.concatenate([p["ct"] for p in parts])
        ok = bool(np.array_equal(lo, ref.kmers_lo) and np.array_equal(ex, ref.exts) and np.array_equal(ct, ref.counts))
if rank == 0:
    nv = max_bytes / (max_ms * 1e-3) / 1e9 if max_ms > 0 else None
    print(json.dumps({"mode": "global_spectrum", "world": world, "reads": int(n_rows), "rounds": st["rounds"], "k": k, "min_coverage": mc,
                      "observations": int(got.n_observations), "survivors": int(tot_surv), "parity_vs_oracle": ok, "wall_ms": dt * 1e3,
                      "obs_per_s": got.n_observations / dt, "nccl_bytes_sent_max_rank": int(max_bytes), "nccl_ms_max_rank": max_ms,
                      "nvlink_GBps_per_dir": nv, "nvlink_frac_of_900": (nv / 900.0) if nv else None,
                      "stage_ms_rank0": {k_: round(v, 2) for k_, v in st.items() if k_.startswith("ms_")}, "level_merges_rank0": st.get("level_merges"),
                      "input": "host numpy (parity run)" if PARITY else "synthesized in HBM, result left in HBM"}))
    assert ok is not False
spectrum.finalize_comm(eng)
rdv.close()
eng.close()
