#!/usr/bin/env python3
"""Summarises an .ncu-rep (read here on the CPU box): headline metrics, SASS opcode mix, stall reasons.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_uniform.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio"]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h:70s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
ops, tot = collections.Counter(), 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stalls = collections.Counter()
for r in data:
    if len(r) < len(hdr):
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
    op = m.group(2).split(".")[0] if m else "?"
    n = int(r[ix["Instructions Executed"]])
    ops[op] += n
    tot += n
    for h in stall_cols:
        stalls[h] += int(r[ix[h]])
print(f"static SASS instructions: {len(data)}   dynamic warp instructions: {tot}")
fp64 = sum(ops[o] for o in ("DFMA", "DMUL", "DADD", "DSETP"))
print(f"FP64 (DFMA+DMUL+DADD+DSETP): {fp64} = {100 * fp64 / tot:.1f} %")
for op, n in ops.most_common(24):
    print(f"  {op:10s} {n:>14d} {100 * n / tot:5.1f} %")
ts = sum(stalls.values())
print("stall samples:", ", ".join(f"{h[6:]} {100 * n / ts:.1f}%" for h, n in stalls.most_common(9)))
