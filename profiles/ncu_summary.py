#!/usr/bin/env python3
"""Summarise an .ncu-rep: headline metrics, SASS opcode mix, stall reasons.
usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [units_per_launch]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
m = {}
for i, h in enumerate(hdr):
    if h in want:
        m[h] = (vals[i], unit[i])
        print(f"{h:62s} {vals[i]:>16s} {unit[i]}")
if units:
    inst = float(m["smsp__inst_executed.sum"][0])
    print(f"warp instructions per unit: {inst / units:.1f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops, stalls, tot = collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    sass = r[ix["Source"]].split()
    n = int(r[ix["Instructions Executed"]] or 0)
    op = (sass[1] if sass[0].startswith("@") else sass[0]).split(".")[0]
    ops[op] += n
    tot += n
    for h in hdr:
        if h.startswith("stall_") and "(Not" not in h:
            stalls[h] += int(r[ix[h]] or 0)
print("\nSASS opcode mix (executed warp instructions)")
for k, v in ops.most_common(18):
    print(f"  {k:10s} {v:13d} {100 * v / tot:5.1f}%")
print("\nwarp stall samples")
ts = sum(stalls.values())
for k, v in stalls.most_common(8):
    print(f"  {k:26s} {v:9d} {100 * v / ts:5.1f}%")
