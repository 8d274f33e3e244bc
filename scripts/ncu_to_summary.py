#!/usr/bin/env python3
"""Writes / updates one entry of profiles/ncu_summary.json from an .ncu-rep captured with `ncu --set full` on the
CURRENT kernel sources (the entry records their sha; bench.py refuses the entry as stale when they change).
usage: python scripts/ncu_to_summary.py <workload key: c2|c3|loss> <file.ncu-rep> <prices per launch> [note] [sha]
(sha: kernel-source sha of the captured build if it is not the working tree, e.g. "round-1")"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_sha  # noqa: E402

key, rep, n_prices = sys.argv[1], sys.argv[2], float(sys.argv[3])
note = sys.argv[4] if len(sys.argv) > 4 else ""
sha = sys.argv[5] if len(sys.argv) > 5 else kernel_source_sha()
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
val = dict(zip(rows[0], rows[2]))
unit = dict(zip(rows[0], rows[1]))


def num(name):
    return float(val[name].replace(",", ""))


def to_bytes(name):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit[name]]
    return num(name) * mult


src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter()
for r in data:
    if len(r) < len(hdr):
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
    ops[m.group(2).split(".")[0] if m else "?"] += int(r[ix["Instructions Executed"]])
fp64 = sum(ops[o] for o in ("DFMA", "DMUL", "DADD", "DSETP"))
flop = 2 * ops["DFMA"] + ops["DMUL"] + ops["DADD"]
time_unit = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit["gpu__time_duration.sum"]]
entry = {
    "kernel": val.get("Kernel Name", "?"),
    "capture": f"{rep} (ncu --set full --clock-control none); text summary beside it in profiles/",
    "kernel_source_sha": sha,
    "gpu_time_ms": num("gpu__time_duration.sum") * time_unit,
    "prices_per_launch": n_prices,
    "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
    "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
    "fp64_pipe_active_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "registers_per_thread": int(num("launch__registers_per_thread")),
    "warp_instructions": sum(ops.values()), "fp64_warp_instructions": fp64,
    "fp64_warp_instructions_per_price": fp64 / n_prices,
    "executed_flop_per_price": flop * 32 / n_prices,
    "opcode_mix_top": {k: v for k, v in ops.most_common(12)},
    "note": note or "fp64 warp instructions = DFMA + DMUL + DADD + DSETP (each holds an FP64 issue slot); executed flop = "
                    "(2 DFMA + DMUL + DADD) x 32 lanes",
}
path = os.path.join(ROOT, "profiles", "ncu_summary.json")
allv = json.load(open(path)) if os.path.exists(path) else {}
allv[key] = entry
json.dump(allv, open(path, "w"), indent=1)
print(json.dumps(entry, indent=1))
