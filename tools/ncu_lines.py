#!/usr/bin/env python3
"""Per-source-line executed-instruction table of one kernel launch from an ncu report captured with --import-source on:
  ncu_lines.py <report.ncu-rep> <kernel regex> <launch-skip> [top N]
Lines are those of csrc/kernels.cu; columns: warp instructions, share, average active threads, stall samples."""
import csv
import io
import subprocess
import sys

rep, rx, skip = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], stdout=subprocess.PIPE, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
lines, fname, h = [], "", None
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        h = r
        ii, it, isamp = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
    elif h and len(r) == len(h) and r[0].strip().isdigit():
        try:
            lines.append((int(r[0]), fname + ": " + r[1].strip()[:100], int(r[ii] or 0), int(r[it] or 0), int(r[isamp] or 0)))
        except ValueError:
            pass
tot_i = sum(l[2] for l in lines) or 1
tot_t = sum(l[3] for l in lines)
tot_s = sum(l[4] for l in lines) or 1
print(f"kernel {rx} launch-skip {skip}: warp instructions {tot_i:,}  threads/instruction {tot_t / tot_i:.2f}  samples {tot_s}")
for ln, src, wi, ti, sm in sorted(lines, key=lambda l: -l[2])[:top]:
    print(f"{ln:5d} {100 * wi / tot_i:5.1f}% inst  {ti / max(wi, 1):5.1f} thr  {100 * sm / tot_s:5.1f}% stall-samples  {src}")
