"""Turn ncu --set full captures (.ncu-rep, read here without a GPU) into profiles/ncu_counters.json: per kernel the DRAM bytes and
warp instructions of ONE launch, the UMI groups that launch covered and the build id of the CUDA sources, so that bench.py can
quote `roofline.traffic` / `roofline.issue` only for the build they were measured on.

usage: python tools/ncu_counters.py UMIS name=REP [name=REP ...]     e.g.  100000 "k_smem_count<1,32,false>"=gpurun_out/r02_smem.ncu-rep
"""
import csv, io, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import kernel_build_id

WANT = {"dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "smsp__inst_executed.sum": "warp_inst",
        "gpu__time_duration.sum": "duration", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_pipe_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct", "launch__registers_per_thread": "registers",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6,
        "nsecond": 1e-9, "second": 1.0}


def read(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, unit, val = rows[0], rows[1], rows[2]
    out = {"kernel_name": val[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""}
    for i, h in enumerate(hdr):
        if h in WANT:
            out[WANT[h]] = float(val[i].replace(",", "")) * UNIT.get(unit[i], 1.0)
    out["dram_bytes"] = out.get("dram_read", 0.0) + out.get("dram_write", 0.0)
    return out


def main():
    umis = int(sys.argv[1])
    kernels = {}
    for a in sys.argv[2:]:
        name, rep = a.split("=", 1)
        k = read(rep)
        k["umis"] = umis
        k["report"] = os.path.basename(rep)
        kernels[name] = k
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_counters.json")
    with open(path, "w") as f:
        json.dump({"build_id": kernel_build_id(), "note": "one launch per kernel, ncu --set full --clock-control none, bench shape (cfg2)",
                   "kernels": kernels}, f, indent=1)
    print(json.dumps(kernels, indent=1))


if __name__ == "__main__":
    main()
