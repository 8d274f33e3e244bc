#!/usr/bin/env python3
"""Which source lines execute the non-FP64 instructions?  Per opcode class, the heaviest CUDA source lines.
usage: python scripts/ncu_ops_by_line.py prof.ncu-rep OP[,OP...] [top]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, ops_want = sys.argv[1], set(sys.argv[2].split(","))
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
fname, hdr, cur = None, None, None
acc = collections.Counter()
src_of = {}
tot = 0
for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 5 and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and len(r) > 5:
        if r[0]:
            cur = (fname, int(r[0])); src_of[cur] = r[1].strip()
        else:
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3].strip())
            op = m.group(2).split(".")[0] if m else "?"
            try:
                n = int(r[hdr["Instructions Executed"]])
            except ValueError:
                continue
            tot += n
            if op in ops_want:
                acc[cur] += n
print(f"{sum(acc.values())} of {tot} warp instructions ({100 * sum(acc.values()) / tot:.1f} %) are {sorted(ops_want)}")
for key, n in acc.most_common(top):
    print(f"{100 * n / tot:5.2f}%  {key[0]}:{key[1]:<4d} {src_of[key][:120]}")
