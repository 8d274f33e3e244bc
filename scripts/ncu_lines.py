#!/usr/bin/env python3
"""Per-CUDA-source-line profile of an .ncu-rep captured with --import-source on (kernels built with -lineinfo):
warp instructions executed and stall samples per source line, heaviest first.
usage: python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
fname, hdr, rows = None, None, []
for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 5 and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and len(r) > 5 and r[0]:
        def num(v):
            try:
                return int(v)
            except ValueError:
                return 0
        rows.append((fname, int(r[0]), r[1].strip(), num(r[hdr["Instructions Executed"]]), num(r[hdr["# Samples"]])))
tot_i = sum(r[3] for r in rows)
tot_s = sum(r[4] for r in rows)
print(f"total warp instructions {tot_i}, samples {tot_s}")
for f, ln, src, ni, ns in sorted(rows, key=lambda r: -r[3])[:top]:
    print(f"{100 * ni / tot_i:5.1f}% inst {100 * ns / tot_s:5.1f}% smpl  {f}:{ln:<4d} {src[:110]}")
