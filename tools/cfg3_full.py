"""BASELINE configs[2] at its stated size (run on the GPU box): 100 000 UMIs with log-uniform reads/UMI in 1..5000 (L = 150, T = 500,
1 % substitutions), k sweep, device-resident; the first UMIS_CHECK groups are compared with the oracle.
usage: cfg3_full.py [n_umis] [k list]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rogtk_b200
from rogtk_b200 import synth
from oracle import dbg_oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [11, 16, 17, 31, 33, 63]
CHECK = 1000
eng = rogtk_b200.Engine()
seed = synth.config_seed(3)
rp = synth.skewed_reads_per_umi(seed, n)
d_data, d_off, d_goff, nbytes, n_reads = eng.synth_reads_device_var(seed, 0, rp, 150, 500, 655)
print(f"cfg3: {n} UMIs, {n_reads} reads, {nbytes / 1e9:.2f} GB, reads/UMI median {int(np.median(rp))} max {int(rp.max())}", flush=True)
# host copy of the first CHECK groups for the oracle
rows = int(rp[:CHECK].sum())
h_data = eng.d2h(d_data, rows * 150, np.uint8)
h_off = np.arange(rows + 1, dtype=np.int64) * 150
h_goff = np.concatenate([[0], np.cumsum(rp[:CHECK])]).astype(np.int64)
# the device generator must equal the numpy one (spot check on the first groups)
ref_data, _, _ = synth.synth_reads(seed, 0, 20, rp[:20], 150, 500)
assert np.array_equal(ref_data, h_data[:len(ref_data)]), "device generator differs from synth.synth_reads"
out = []
for k in ks:
    best = None
    for it in range(3):
        db = eng.count_prune_device(d_data, nbytes, d_off, n_reads, d_goff, n, k, 2)
        r = (db.ms_total, db.ms_hot_kernel, db.n_observations, db.n_kmers, db.k_eff, db.n_multipass_groups, db.n_spilled_groups)
        if it == 2:
            got = db.to_host() if False else None
            # parity of the first CHECK groups: pull only their slice
            off = eng.d2h(db._c.kmer_offsets, (CHECK + 1) * 8, np.int64)
            a = int(off[CHECK])
            lo = eng.d2h(db._c.kmers_lo, a * 8, np.uint64)
            ex = eng.d2h(db._c.exts, a, np.uint8)
            ct = eng.d2h(db._c.counts, a * 2, np.uint16)
            hi = eng.d2h(db._c.kmers_hi, a * 8, np.uint64) if db.k_eff > 32 else None
        db.free()
        if best is None or r[0] < best[0]:
            best = r
    ref = dbg_oracle.count_prune_batch(h_data, h_off, h_goff, k, 2)
    ok = (np.array_equal(off, ref.kmer_offsets) and np.array_equal(lo, ref.kmers_lo) and np.array_equal(ex, ref.exts) and np.array_equal(ct, ref.counts)
          and (hi is None or np.array_equal(hi, ref.kmers_hi)))
    ms, msk, nobs, nk, ke, nm, ns = best
    rec = dict(config="cfg3", umis=n, reads=n_reads, k=k, k_eff=ke, observations=nobs, survivors=nk, device_ms=ms, first_tier_kernel_ms=msk,
               multipass_groups=nm, hbm_table_groups=ns, obs_per_s=nobs / (ms * 1e-3), parity_first_1000_groups=bool(ok))
    print(json.dumps(rec), flush=True)
    out.append(rec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/cfg3_full.json", "w"), indent=1)
