"""Summarise an .ncu-rep (read here, no GPU needed): key metrics + SASS hot regions.  usage: ncu_summary.py REP [lo hi]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit, val = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum", "lts__t_sectors_op_atom.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram__bytes_write.sum.pct_of_peak_sustained_elapsed", "dram__bytes.sum.per_second",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__maximum_warps_per_active_cycle_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
for i, h in enumerate(hdr):
    if h in want or ("issue_stalled" in h and "per_issue_active" in h and "average" in h):
        print(f"{h:95s} {unit[i]:12s} {val[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
ts = sum(int(r[ix["# Samples"]]) for r in data)
print("total warp-inst", tot, "samples", ts, "sass lines", len(data))
for b in range(0, len(data), 100):
    ins = sum(int(r[ix["Instructions Executed"]]) for r in data[b:b + 100])
    sm = sum(int(r[ix["# Samples"]]) for r in data[b:b + 100])
    if ins > tot * 0.01 or sm > ts * 0.01:
        print(f"  sass[{b}:{b+100}] inst {100*ins/tot:5.1f}%  samples {100*sm/ts:5.1f}%")
if len(sys.argv) > 3:
    for i in range(int(sys.argv[2]), int(sys.argv[3])):
        r = data[i]
        print(i, r[ix["Source"]].strip()[:62].ljust(62), r[ix["Instructions Executed"]].rjust(10), r[ix["Avg. Threads Executed"]].rjust(3), r[ix["# Samples"]].rjust(6))
