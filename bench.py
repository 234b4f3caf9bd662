#!/usr/bin/env python
"""bench.py -- per-UMI de Bruijn k-mer count/prune throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--umis U]

A "step" is one pass of the hot path (rogtk_gpu_count_prune_batch) over one batch of synthetic reads.  At N=1 the
workload is BASELINE.json configs[1]: 1M UMIs x 50 reads x 300 bp, k=31 (K_eff=32), min_coverage=2, generated directly in
HBM (SURVEY 8d generator).  N>1: one process per GPU (torchrun), every rank processes its own 1M-UMI shard (UMI groups are
independent: no collective on the data path) -> weak scaling.

  value : k-mer observations / s over all ranks, inputs resident in HBM, outputs left in HBM, device pipeline time from
          CUDA events on the library's stream (max over ranks).
  e2e   : the same metric through the C ABI with HOST (pinned) Arrow buffers in and host arrays out: H2D + D2H inside the
          timed region (wall clock around the synchronous call).
  roofline : algorithmic HBM bytes of the dominant kernel (k_smem_count) / its CUDA-event duration vs MEASURED_PEAKS.json.
  cpu_baseline : the C restatement of the reference algorithm (oracle/, OpenMP over groups) on a bounded sample.

torch is used only for torch.distributed plumbing (barrier / max over ranks); the library allocates its own device memory.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import ctypes

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG2 = dict(name="cfg2", n_umis=1_000_000, reads=50, read_len=300, template=1000, k=31, min_cov=2, err64k=655, config_id=2)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None
        self.t_mark = 0.0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        """start of the timed region: earlier samples are discarded"""
        self.t_mark = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ts, ln in self.lines:
            if ts < self.t_mark:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples in the upper half of the observed power range
        p_thr = (max(power) + min(power)) / 2
        load = [s for s, p in zip(sm, power) if p >= p_thr] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def kernel_build_id() -> str:
    """sha1 over the kernel sources (kernels_*.cuh and the device headers they share; the host-side files -- C ABI, drivers, path
    finding, plugin -- do not change a kernel): ncu-derived counters are only quoted for the build they were captured on"""
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, "rogtk_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.startswith("kernels_") or f in ("common.cuh", "scan.cuh"):
            with open(os.path.join(d, f), "rb") as fh:
                h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


def ncu_counters():
    """profiles/ncu_counters.json (written by tools/ncu_counters.py from an ncu --set full capture): per kernel the DRAM bytes and
    warp instructions of ONE launch over `umis` UMI groups of the bench shape, and the build id of the sources it ran on."""
    p = os.path.join(ROOT, "profiles", "ncu_counters.json")
    if not os.path.exists(p):
        return {}, None
    with open(p) as f:
        d = json.load(f)
    return d.get("kernels", {}), d.get("build_id")


def algorithmic_bytes(n_groups, n_reads, n_bytes, n_kmers, key_bytes=8):
    """read every input byte + offsets once, write every survivor once.
    B_in = ASCII bytes + 8*(reads+1) per group ; B_out = S*(W+1+2) + 8 per group."""
    return n_bytes + 8 * (n_reads + n_groups) + n_kmers * (key_bytes + 3) + 8 * n_groups


def cpu_baseline(cfg, host=None, target_seconds=15.0, threads=0):
    """Time the oracle (C restatement of the reference algorithm, one group per OpenMP task) on a bounded sample of the
    workload: a prefix of the host copy of the bench input when there is one (host = (data, offsets, group_offsets) numpy
    views), else a numpy-generated prefix."""
    from oracle import dbg_oracle
    from rogtk_b200 import synth
    seed = synth.config_seed(cfg["config_id"])
    if threads <= 0:
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else dbg_oracle.max_threads()
    cores = threads
    R, L = cfg["reads"], cfg["read_len"]

    def run(n):
        if host is not None:
            data, offs, goff = host[0][:n * R * L], host[1][:n * R + 1], host[2][:n + 1]
        else:
            data, offs, goff = synth.synth_reads_chunked(seed, 0, n, R, L, cfg["template"], cfg["err64k"])
        t0 = time.perf_counter()
        b = dbg_oracle.count_prune_batch(data, offs, goff, cfg["k"], cfg["min_cov"], n_threads=cores)
        return time.perf_counter() - t0, b.n_observations

    n0 = min(cfg["n_umis"], max(4096, 512 * cores))
    t, obs = run(n0)
    rate = obs / max(t, 1e-9)
    per_umi = obs / n0
    n = int(min(cfg["n_umis"], max(n0, rate * target_seconds / per_umi)))
    if host is None:
        n = min(n, 60000)   # bound generation time / memory when the sample has to be generated on the host
    t, obs = run(n)
    # one thread on a ~2 s sample (SURVEY 8d asks for the single-thread rate beside the all-cores one)
    n1 = int(min(n, max(256, rate / max(cores, 1) * 2.0 / per_umi)))
    all_cores = cores
    cores = 1
    t1, obs1 = run(n1)
    cores = all_cores
    return {"value": obs / t, "unit": "obs/s", "cores": cores, "kind": "port",
            "sample": f"first {n} UMIs of {cfg['name']} ({obs} observations, {t:.2f} s); C restatement of the reference "
                      f"algorithm (oracle/dbg_oracle.c), OpenMP dynamic schedule over groups",
            "umis_per_s": n / t,
            "single_thread": {"value": obs1 / t1, "unit": "obs/s", "sample": f"first {n1} UMIs ({t1:.2f} s)"}}


def eng_walk_bits() -> int:
    """the first-tier counting kernel the library will launch (ROGTK_WALK_RL, read by rogtk_gpu_create)"""
    try:
        v = int(os.environ.get("ROGTK_WALK_RL", "0"))
    except ValueError:
        v = 0
    return 0 if v <= 0 else min(max(v, 3), 6)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--umis", type=int, default=0, help="override UMIs per GPU (debug; the default is the BASELINE config)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = dict(CFG2)
    if args.umis > 0:
        cfg["n_umis"] = args.umis
    workload = (f"{cfg['name']}: {cfg['n_umis']} UMIs x {cfg['reads']} reads x {cfg['read_len']} bp per GPU, k={cfg['k']} "
                f"(K_eff=32), min_coverage={cfg['min_cov']}, template {cfg['template']} bp, 1% substitutions")
    base = {"metric": "kmer_observations_per_s", "unit": "obs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic (SplitMix64 generator of SURVEY 8d, generated in HBM)"}

    # ------------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        # The Rust reference cannot be built here (no cargo; debruijn/boomphf crates not vendored): the arm times the
        # oracle port of its algorithm on the host cores, each step a bounded sample of the same workload.
        from oracle import dbg_oracle
        from rogtk_b200 import synth
        # every host core the process may use (torchrun exports OMP_NUM_THREADS=1, which is not what this arm measures)
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        n = min(cfg["n_umis"], 65536, max(8192, 2048 * cores))
        data, offs, goff = synth.synth_reads_chunked(synth.config_seed(cfg["config_id"]), 0, n, cfg["reads"], cfg["read_len"],
                                                     cfg["template"], cfg["err64k"])
        obs = 0
        for _ in range(max(1, min(args.warmup, 1))):
            obs = dbg_oracle.count_prune_batch(data, offs, goff, cfg["k"], cfg["min_cov"], n_threads=cores).n_observations
        t0 = time.perf_counter()
        for _ in range(args.steps):
            dbg_oracle.count_prune_batch(data, offs, goff, cfg["k"], cfg["min_cov"], n_threads=cores)
        dt = (time.perf_counter() - t0) / args.steps
        val = obs / dt
        line = dict(base)
        line.update({"impl": "reference", "value": val, "ms_per_step": dt * 1e3, "umi_groups_per_s": n / dt,
                     "config": {"workload": workload, "step": f"bounded sample: first {n} UMIs per step"},
                     "cpu_baseline": {"value": val, "unit": "obs/s", "cores": cores, "kind": "port",
                                      "sample": f"first {n} UMIs of {cfg['name']} per step ({obs} observations)"},
                     "e2e": {"value": val, "unit": "obs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "gpu_launches": 0})
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------------------------------ our arm
    dist = None
    if world > 1:
        # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION; the contract is ONE JSON line on stdout
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # NCCL writes its version banner to stdout when the communicator is created; stdout must carry ONE JSON line, so the
        # process-level stdout is pointed at stderr while the communicator comes up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    import rogtk_b200
    from rogtk_b200 import synth
    eng = rogtk_b200.Engine(device=local_rank)
    seed = synth.config_seed(cfg["config_id"])
    n_umis, R, L = cfg["n_umis"], cfg["reads"], cfg["read_len"]
    n_reads = n_umis * R
    # each rank owns a disjoint range of UMI indices (its hash shard, materialised as a contiguous range)
    # each rank owns the UMIs its hash names: the first n_umis ids u >= 0 with rogtk_gpu_shard_of(u, world) == rank (what a host
    # that shards a frame by rogtk_gpu_shard_of hands to this GPU); one rank: ids 0 .. n_umis - 1
    shard_ids_ = None
    if world > 1:
        from rogtk_b200 import sharding
        pop = np.arange(int(world * n_umis * 1.02) + 4096, dtype=np.int64)
        shard_ids_ = pop[sharding.shard_ids(pop, world) == rank][:n_umis]
        assert len(shard_ids_) == n_umis
        del pop
    d_data, d_off, d_goff, n_bytes = eng.synth_reads_device(seed, rank * n_umis, n_umis, R, L, cfg["template"], cfg["err64k"], umi_ids=shard_ids_)

    stage_acc = {}

    def step(flags=0, acc=None):
        db = eng.count_prune_device(d_data, n_bytes, d_off, n_reads, d_goff, n_umis, cfg["k"], cfg["min_cov"], flags=flags)
        r = (db.ms_total, db.ms_hot_kernel, db.n_observations, db.n_kmers, db.n_spilled_groups)
        if acc is not None:
            for k_, v_ in db.stage_ms.items():
                acc[k_] = acc.get(k_, 0.0) + v_
            acc["n_edges"] = db.n_edges
        db.free()
        return r

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()   # nvidia-smi needs ~1 s to produce its first line: start it before the warm-up
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = eng.launch_count
    sampler.mark()
    t0 = time.perf_counter()
    ms_dev, ms_hot = 0.0, 0.0
    for _ in range(args.steps):
        ms_total, ms_k, n_obs, n_kmers, n_spill = step(acc=stage_acc)
        ms_dev += ms_total
        ms_hot += ms_k
    wall = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count - launches0
    ms_dev_max = max_over_ranks(ms_dev) / args.steps
    wall_max = max_over_ranks(wall) / args.steps * 1e3
    ms_hot_avg = ms_hot / args.steps
    obs_all = sum_over_ranks(float(n_obs))
    groups_all = sum_over_ranks(float(n_umis))
    value = obs_all / (ms_dev_max * 1e-3)

    peak, peak_src = load_peaks()
    alg = algorithmic_bytes(n_umis, n_reads, n_bytes, n_kmers)
    achieved = alg / (ms_hot_avg * 1e-3) / 1e9 if ms_hot_avg > 0 else 0.0
    # ncu-derived counters (DRAM bytes, warp instructions) come from profiles/ncu_counters.json and are quoted only when that capture
    # ran on the sources this library was built from; otherwise they are null (stale) rather than silently wrong.
    counters, cap_build = ncu_counters()
    build_id = kernel_build_id()
    fresh = cap_build == build_id

    def per_launch(kernel, field):
        c = counters.get(kernel)
        if not (fresh and c and cfg["name"] == "cfg2"):
            return None
        return c[field] / c["umis"] * n_umis

    hot_name = "k_smem_count<1,32,false>" if eng_walk_bits() == 0 else "k_walk_count<1>"
    traffic = per_launch(hot_name, "dram_bytes")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": hot_name, "kernel_ms": ms_hot_avg, "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                "pipeline_frac": alg / (ms_dev_max * 1e-3) / 1e9 / peak,
                "counters": {"source": "profiles/ncu_counters.json", "capture_build": cap_build, "this_build": build_id,
                             "fresh": fresh}}
    wi = per_launch(hot_name, "warp_inst")
    if wi is not None and ms_hot_avg > 0:
        # What actually limits the kernel: warp instructions issued, against one per scheduler per clock
        # (SMs x 4 x the SM clock sampled during the timed region).
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        inst_rate = wi / (ms_hot_avg * 1e-3)
        inst_peak = 148 * 4 * sm_mhz * 1e6
        roofline["issue"] = {"achieved_warp_inst_per_s": inst_rate, "peak_warp_inst_per_s": inst_peak, "frac": inst_rate / inst_peak,
                             "warp_inst_per_umi": wi / n_umis}
    # every kernel group of the step, timed live with CUDA events on the library's stream (averages over the timed steps)
    st_ms = {k_: v_ / args.steps for k_, v_ in stage_acc.items() if k_ != "n_edges"}
    sort_bytes = n_kmers * 22 + 16 * n_umis          # pool read + dense write per survivor, offsets per group
    kernels = [{"kernel": hot_name, "bound": "hbm", "ms": ms_hot_avg, "algorithmic_bytes": alg,
                "achieved": achieved, "peak": peak, "frac": achieved / peak, "traffic": traffic},
               {"kernel": "k_sort_small (+mid/big tiers)", "bound": "hbm", "ms": st_ms.get("sort"), "algorithmic_bytes": sort_bytes,
                "achieved": sort_bytes / (st_ms["sort"] * 1e-3) / 1e9 if st_ms.get("sort") else None, "peak": peak,
                "frac": sort_bytes / (st_ms["sort"] * 1e-3) / 1e9 / peak if st_ms.get("sort") else None,
                "traffic": per_launch("k_sort_small", "dram_bytes")}]
    roofline["kernels"] = kernels

    # ---- the same step with the successor CSR and the node statistics switched on (north_star: "CSR adjacency emit") ------------
    value_with_csr = None
    if rank == 0 or world > 1:
        from rogtk_b200 import _ffi as _f
        csr_acc = {}
        step(flags=_f.EMIT_CSR | _f.EMIT_STATS)
        ms_csr_total = 0.0
        n_csr_steps = max(1, min(args.steps, 3))
        for _ in range(n_csr_steps):
            ms_csr_total += step(flags=_f.EMIT_CSR | _f.EMIT_STATS, acc=csr_acc)[0]
        ms_csr_step = max_over_ranks(ms_csr_total / n_csr_steps)
        n_edges = csr_acc.get("n_edges", 0)
        csr_ms = csr_acc.get("csr", 0.0) / n_csr_steps
        stats_ms = csr_acc.get("stats", 0.0) / n_csr_steps
        csr_bytes = n_kmers * (8 + 1) * 2 + n_kmers * 8 + n_edges * 4      # keys + exts read in both passes, row_ptr + col written
        stats_bytes = n_kmers + 12 * n_umis
        value_with_csr = {"value": obs_all / (ms_csr_step * 1e-3), "unit": "obs/s", "ms_per_step": ms_csr_step, "steps": n_csr_steps,
                          "flags": "ROGTK_EMIT_CSR | ROGTK_EMIT_STATS", "edges_per_step": int(n_edges)}
        kernels.append({"kernel": "k_csr (degree pass + scan + fill)", "bound": "hbm", "ms": csr_ms, "algorithmic_bytes": csr_bytes,
                        "achieved": csr_bytes / (csr_ms * 1e-3) / 1e9 if csr_ms else None, "peak": peak,
                        "frac": csr_bytes / (csr_ms * 1e-3) / 1e9 / peak if csr_ms else None, "traffic": None})
        kernels.append({"kernel": "k_stats", "bound": "hbm", "ms": stats_ms, "algorithmic_bytes": stats_bytes,
                        "achieved": stats_bytes / (stats_ms * 1e-3) / 1e9 if stats_ms else None, "peak": peak,
                        "frac": stats_bytes / (stats_ms * 1e-3) / 1e9 / peak if stats_ms else None, "traffic": None})

    # ---- next row f-2 (outside the timed region): unitig compaction + the plugin's contig on the same device-resident batch ------
    next_rows = None
    if rank == 0 and args.gpus == 1:
        try:
            db = eng.count_prune_device(d_data, n_bytes, d_off, n_reads, d_goff, n_umis, cfg["k"], cfg["min_cov"])
            nu, nb, ms_u = eng.compress_unitigs(db, 200, sizes_only=True)
            nu, nb, ms_u = eng.compress_unitigs(db, 200, sizes_only=True)
            db.free()
            next_rows = {"unitig_compaction": {"ms": ms_u, "unitigs": nu, "contig_bases": nb, "min_length": 200,
                                               "api": "rogtk_gpu_compress_unitigs, device-resident in and out"}}
        except Exception as exc:   # reported, never fatal for the headline line
            next_rows = {"unitig_compaction": {"error": str(exc)[:200]}}

    # ---- next rows outside the timed region: sweep / optimize drivers (count once per K_eff vs one recount per probe), and for
    #      N > 1 the global k-mer spectrum with the library's own NCCL exchange over NVLink --------------------------------------------
    if rank == 0 and args.gpus == 1 and not args.no_e2e:
        try:
            n_s = min(n_umis, 20000)
            sd = eng.d2h(d_data, n_s * R * L, np.uint8)
            so = eng.d2h(d_off, (n_s * R + 1) * 8, np.int64)
            sg = eng.d2h(d_goff, (n_s + 1) * 8, np.int64)
            sweep_kw = dict(k_start=25, k_end=40, k_step=5, cov_start=1, cov_end=6, cov_step=1, method="compression")
            eng.sweep_assembly_params(sd, so, sg, **sweep_kw)   # warm-up: the first call of a new mix of sizes pays ~60 cudaMallocs
            eng.assemble_sequences(sd, so, sg, k=31, min_coverage=2, method="compression")
            t0 = time.perf_counter()
            sw = eng.sweep_assembly_params(sd, so, sg, **sweep_kw)
            dt_sw = time.perf_counter() - t0
            cells = sw["k"].shape[1]
            t0 = time.perf_counter()
            one = eng.assemble_sequences(sd, so, sg, k=31, min_coverage=2, method="compression")
            dt_one = time.perf_counter() - t0
            next_rows = dict(next_rows or {})
            next_rows["sweep"] = {"umis": n_s, "cells": int(cells), "count_passes": sw["n_count_passes"], "wall_ms": dt_sw * 1e3,
                                  "one_cell_wall_ms": dt_one * 1e3, "speedup_vs_recount_per_cell": dt_one * cells / dt_sw,
                                  "api": "rogtk_gpu_sweep_batch: k 25..40 step 5 x min_coverage 1..6, compression"}
        except Exception as exc:
            next_rows = dict(next_rows or {})
            next_rows["sweep"] = {"error": str(exc)[:200]}
        try:   # the reference's default family of methods: count + CSR + path search per group, all on the device; reads resident in HBM in,
            #    contigs on the host out (first 200 000 groups of the batch)
            from rogtk_b200 import _ffi as _f2
            n_p = min(n_umis, 200000)
            dptr_p = (ctypes.c_void_p * 1)(d_data)
            dlen_p = (ctypes.c_int64 * 1)(n_p * R * L)
            col_p = _f2.StrCol(layout=_f2.UTF8_I64, n_buffers=1, n_rows=n_p * R, validity=None, validity_bit_offset=0, offsets_or_views=d_off,
                               data=dptr_p, data_len=dlen_p)
            ap_p = _f2.AssemblyParamsC(k=cfg["k"], min_coverage=cfg["min_cov"], method=_f2.METHODS["shortest_path_auto"], min_length=0, only_largest=1)
            for _ in range(2):   # (the first call sizes the cached blocks)
                so_p = _f2.StringBatchC()
                t0 = time.perf_counter()
                eng._check(eng._lib.rogtk_gpu_assemble_batch(eng._ctx, ctypes.byref(col_p), ctypes.c_void_p(d_goff), n_p, ctypes.byref(ap_p),
                                                             _f2.INPUT_ON_DEVICE, ctypes.byref(so_p)))
                dt_sp = time.perf_counter() - t0
                offs_p = np.ctypeslib.as_array(ctypes.cast(so_p.offsets, ctypes.POINTER(ctypes.c_int64)), shape=(n_p + 1,)).copy()
                ms_dev_p, ms_host_p = float(so_p.ms_device), float(so_p.ms_host)
                eng._lib.rogtk_gpu_free_strings(eng._ctx, ctypes.byref(so_p))
            lens = np.diff(offs_p)
            next_rows = dict(next_rows or {})
            next_rows["shortest_path_auto"] = {"umis": n_p, "wall_ms": dt_sp * 1e3, "ms_device": ms_dev_p, "ms_host": ms_host_p,
                                               "groups_per_s": n_p / dt_sp, "groups_with_contig": int((lens > 0).sum()),
                                               "mean_contig_bp": float(lens[lens > 0].mean()) if (lens > 0).any() else 0.0,
                                               "api": "rogtk_gpu_assemble_batch(method=shortest_path_auto, ROGTK_INPUT_ON_DEVICE): count + CSR + "
                                                      "petgraph-exact search (one thread per group, csrc/pathfind_device.cuh) on the GPU"}
        except Exception as exc:
            next_rows = dict(next_rows or {})
            next_rows["shortest_path_auto"] = {"error": str(exc)[:200]}
    if world > 1:
        spec = None
        try:
            from rogtk_b200 import spectrum
            port = int(os.environ.get("MASTER_PORT", "29500")) + 23
            rdv = spectrum.Rendezvous(rank, world, "127.0.0.1", port)
            spectrum.init_comm(eng, rdv)
            n_tmpl, per, Ls, Ts = 4000, 50, 150, 500
            sdata, soffs, _ = synth.synth_reads(synth.config_seed(5), 0, n_tmpl, per, Ls, Ts)
            n_rows_s = len(soffs) - 1
            rows = np.arange(rank, n_rows_s, world)
            dd = sdata.reshape(n_rows_s, Ls)[rows].reshape(-1)
            oo = np.arange(len(rows) + 1, dtype=np.int64) * Ls
            st_ = {}
            for _ in range(2):
                rdv.barrier()
                t0 = time.perf_counter()
                got = spectrum.global_spectrum(eng, dd, oo, 31, 2, rdv=rdv, stats_out=st_)
                rdv.barrier()
                dt_s = time.perf_counter() - t0
            np.savez(f"/tmp/rogtk_spec_{rank}.npz", lo=got.kmers_lo, ex=got.exts, ct=got.counts)
            rdv.barrier()
            mx_b = rdv.allreduce(st_["bytes_sent"], "max")
            mx_ms = rdv.allreduce(int(st_["ms_exchange"] * 1000), "max") / 1000.0
            if rank == 0:
                from oracle import dbg_oracle
                ref = dbg_oracle.count_prune_batch(sdata, soffs, np.array([0, n_rows_s]), 31, 2)
                parts = [np.load(f"/tmp/rogtk_spec_{r}.npz") for r in range(world)]
                ok = (np.array_equal(np.concatenate([p_["lo"] for p_ in parts]), ref.kmers_lo)
                      and np.array_equal(np.concatenate([p_["ex"] for p_ in parts]), ref.exts)
                      and np.array_equal(np.concatenate([p_["ct"] for p_ in parts]), ref.counts))
                gbps = mx_b / (mx_ms * 1e-3) / 1e9 if mx_ms > 0 else None
                spec = {"ranks": world, "reads": int(n_rows_s), "observations": int(got.n_observations), "parity_vs_oracle": bool(ok),
                        "wall_ms": dt_s * 1e3, "obs_per_s": got.n_observations / dt_s, "nccl_bytes_sent_max_rank": int(mx_b),
                        "nccl_ms_max_rank": mx_ms, "nvlink_GBps_per_dir": gbps, "nvlink_frac_of_900": gbps / 900.0 if gbps else None,
                        "api": "rogtk_gpu_spectrum_exchange (grouped ncclSend/ncclRecv inside librogtk_gpu.so)"}
            spectrum.finalize_comm(eng)
            rdv.close()
        except Exception as exc:
            spec = {"error": str(exc)[:300]}
        if rank == 0:
            next_rows = dict(next_rows or {})
            next_rows["global_spectrum"] = spec

    # ---- e2e: host (pinned) Arrow buffers -> C ABI -> host arrays; H2D + D2H inside the timed region -------------------
    e2e = None
    cpu = None
    if not args.no_e2e:
        # host memory is shared by all ranks of the node: the e2e leg takes the first e_umis UMIs of the rank's shard, as many
        # as fit in 40 % of the available host memory divided by the ranks (pinned input + pinned result, ~27 KB per UMI)
        try:
            with open("/proc/meminfo") as f:
                avail = next(int(l.split()[1]) * 1024 for l in f if l.startswith("MemAvailable"))
        except Exception:
            avail = 64 << 30
        per_umi = R * L + 8 * R + 8 + 12 * 1100
        e_umis = int(min(n_umis, max(1000, 0.4 * avail / world / per_umi)))
        full_umis, full_reads, full_bytes = n_umis, n_reads, n_bytes
        n_umis, n_reads, n_bytes = e_umis, e_umis * R, e_umis * R * L
        h_data = eng.pinned_alloc(n_bytes + 16)
        h_off = eng.pinned_alloc((n_reads + 1) * 8)
        h_goff = eng.pinned_alloc((n_umis + 1) * 8)
        lib, ctx = eng._lib, eng._ctx
        for dst, src, nb in ((h_data, d_data, n_bytes), (h_off, d_off, (n_reads + 1) * 8), (h_goff, d_goff, (n_umis + 1) * 8)):
            eng._check(lib.rogtk_gpu_memcpy_d2h(ctx, ctypes.c_void_p(dst), ctypes.c_void_p(src), nb))
        from rogtk_b200 import _ffi
        dptr = (ctypes.c_void_p * 1)(h_data)
        dlen = (ctypes.c_int64 * 1)(n_bytes)
        col = _ffi.StrCol(layout=_ffi.UTF8_I64, n_buffers=1, n_rows=n_reads, validity=None, validity_bit_offset=0,
                          offsets_or_views=h_off, data=dptr, data_len=dlen)

        def e2e_step():
            out = _ffi.GraphBatchC()
            eng._check(lib.rogtk_gpu_count_prune_batch(ctx, ctypes.byref(col), ctypes.c_void_p(h_goff), n_umis, cfg["k"],
                                                       cfg["min_cov"], 0, ctypes.byref(out)))
            nk = int(out.n_kmers)
            chk = int(ctypes.cast(out.kmer_offsets, ctypes.POINTER(ctypes.c_int64))[n_umis])   # read the result on the host
            assert chk == nk
            lib.rogtk_gpu_free_batch(ctx, ctypes.byref(out))
            return nk

        e2e_step()   # warm-up: pinned result buffers get cached in the context
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            nk = e2e_step()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        barrier()
        dt = max_over_ranks(dt)
        h2d = n_bytes + (n_reads + 1) * 8 + (n_umis + 1) * 8
        d2h = nk * 11 + (n_umis + 1) * 8
        e2e_obs = sum_over_ranks(float(n_reads * (L - 31)))
        e2e = {"value": e2e_obs / dt, "unit": "obs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": dt * 1e3, "steps": args.e2e_steps, "umis_per_rank": n_umis,
               "api": "rogtk_gpu_count_prune_batch (C ABI via ctypes), pinned host Arrow buffers in, pinned host arrays out"}

        def view(ptr, n, dt_):
            ct = np.ctypeslib.as_ctypes_type(np.dtype(dt_))
            return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ct)), shape=(n,))

        # ---- the same call on what a Polars / pyarrow caller really holds: PAGEABLE buffers, and the Utf8View layout -----------------
        variants = [{"name": "pinned LargeUtf8", "ms_per_step": dt * 1e3, "value": e2e_obs / dt}]
        if world == 1:
            pv = (view(h_data, n_bytes, np.uint8), view(h_off, n_reads + 1, np.int64), view(h_goff, n_umis + 1, np.int64))
            p_data, p_off, p_goff = (np.array(x, copy=True) for x in pv)     # ordinary (pageable) allocations

            def timed(fn, reps=1):
                fn()
                t0_ = time.perf_counter()
                for _ in range(reps):
                    fn()
                return (time.perf_counter() - t0_) / reps

            dptr2 = (ctypes.c_void_p * 1)(p_data.ctypes.data)
            col2 = _ffi.StrCol(layout=_ffi.UTF8_I64, n_buffers=1, n_rows=n_reads, validity=None, validity_bit_offset=0,
                               offsets_or_views=p_off.ctypes.data, data=dptr2, data_len=dlen)

            def pageable_step():
                out = _ffi.GraphBatchC()
                eng._check(lib.rogtk_gpu_count_prune_batch(ctx, ctypes.byref(col2), ctypes.c_void_p(p_goff.ctypes.data), n_umis, cfg["k"],
                                                           cfg["min_cov"], 0, ctypes.byref(out)))
                lib.rogtk_gpu_free_batch(ctx, ctypes.byref(out))

            dt_p = timed(pageable_step, args.e2e_steps)
            variants.append({"name": "pageable LargeUtf8", "ms_per_step": dt_p * 1e3, "value": e2e_obs / dt_p})
            # Utf8View: 16-byte views + data buffers of < 2 GiB (Arrow's variadic buffers), every read out of line (300 B > 12)
            chunk = (1 << 31) // L * L
            o = p_off[:-1]
            views = np.zeros((n_reads, 4), dtype=np.uint32)
            views[:, 0] = L
            views[:, 1] = p_data[:4].view(np.uint32)[0] if n_bytes >= 4 else 0        # prefix bytes (not read by the library)
            views[:, 2] = (o // chunk).astype(np.uint32)
            views[:, 3] = (o % chunk).astype(np.uint32)
            nb_ = int((n_bytes + chunk - 1) // chunk)
            bufs = [p_data[i * chunk:min(n_bytes, (i + 1) * chunk)] for i in range(nb_)]
            bptr = (ctypes.c_void_p * nb_)(*[b_.ctypes.data for b_ in bufs])
            blen = (ctypes.c_int64 * nb_)(*[b_.size for b_ in bufs])
            col3 = _ffi.StrCol(layout=_ffi.UTF8_VIEW, n_buffers=nb_, n_rows=n_reads, validity=None, validity_bit_offset=0,
                               offsets_or_views=views.ctypes.data, data=bptr, data_len=blen)

            def view_step():
                out = _ffi.GraphBatchC()
                eng._check(lib.rogtk_gpu_count_prune_batch(ctx, ctypes.byref(col3), ctypes.c_void_p(p_goff.ctypes.data), n_umis, cfg["k"],
                                                           cfg["min_cov"], 0, ctypes.byref(out)))
                assert int(out.n_kmers) == nk
                lib.rogtk_gpu_free_batch(ctx, ctypes.byref(out))

            dt_v = timed(view_step, 1)
            variants.append({"name": "pageable Utf8View", "ms_per_step": dt_v * 1e3, "value": e2e_obs / dt_v, "data_buffers": nb_})
            # the plugin-level call: reads in, contigs out (Compression; about 1.2 KB per UMI come back instead of 10.5 KB of nodes)
            ap = _ffi.AssemblyParamsC(k=cfg["k"], min_coverage=cfg["min_cov"], method=_ffi.METHODS["compression"], min_length=200, only_largest=1)

            def asm_step():   # through the C ABI like e2e_step (the Python wrapper would add a copy of the result into numpy arrays)
                so = _ffi.StringBatchC()
                eng._check(lib.rogtk_gpu_assemble_batch(ctx, ctypes.byref(col), ctypes.c_void_p(h_goff), n_umis, ctypes.byref(ap), 0, ctypes.byref(so)))
                nb_out = int(ctypes.cast(so.offsets, ctypes.POINTER(ctypes.c_int64))[n_umis])   # read the result on the host
                assert nb_out == int(so.n_bytes)
                r = (nb_out + (n_umis + 1) * 8 + n_umis, float(so.ms_device))
                lib.rogtk_gpu_free_strings(ctx, ctypes.byref(so))
                return r

            asm_step()   # warm-up: the page-locked result blocks get cached in the context
            t0_ = time.perf_counter()
            for _ in range(args.e2e_steps):
                d2h_a, ms_dev_a = asm_step()
            dt_a = (time.perf_counter() - t0_) / args.e2e_steps
            variants.append({"name": "assemble_sequences(compression) pinned in, contigs out", "ms_per_step": dt_a * 1e3, "value": e2e_obs / dt_a,
                             "d2h_bytes": d2h_a, "ms_device": ms_dev_a, "api": "rogtk_gpu_assemble_batch (C ABI via ctypes)"})
            del p_data, p_off, p_goff, views, bufs
        e2e["variants"] = variants
        host_views = None
        if rank == 0 and args.gpus == 1 and not args.no_cpu:
            host_views = (view(h_data, n_bytes, np.uint8), view(h_off, n_reads + 1, np.int64), view(h_goff, n_umis + 1, np.int64))
            cpu = cpu_baseline(cfg, host=host_views)
        for p in (h_data, h_off, h_goff):
            eng.pinned_free(p)
        n_umis, n_reads, n_bytes = full_umis, full_reads, full_bytes

    if cpu is None and rank == 0 and args.gpus == 1 and not args.no_cpu:
        cpu = cpu_baseline(cfg)

    if rank == 0:
        line = dict(base)
        line.update({
            "value": value, "ms_per_step": ms_dev_max, "wall_ms_per_step": wall_max, "umi_groups_per_s": groups_all / (ms_dev_max * 1e-3),
            "config": {"workload": workload, "parallelism": f"umi-sharded x{world} by rogtk_gpu_shard_of(umi_id) (splitmix64 mod ranks), no collective",
                       "l2": "inputs (15 GB/GPU) far larger than the 126 MB L2; no flush needed",
                       "timing": "CUDA events on the library stream around the whole device pipeline, max over ranks"},
            "roofline": roofline, "value_with_csr": value_with_csr, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "device_cache": eng.cache_stats(),
            "clocks": clocks,
            "survivors_per_step": int(n_kmers), "spilled_groups": int(n_spill), "observations_per_step": int(obs_all),
            "next_rows": next_rows})
        print(json.dumps(line))
    for p in (d_data, d_off, d_goff):
        eng.device_free(p)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
