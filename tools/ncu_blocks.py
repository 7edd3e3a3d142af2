#!/usr/bin/env python
"""Basic-block view of the first kernel in an .ncu-rep: stretches of SASS with the same execution count, their
share of all executed warp-instructions and of the stall samples.  usage: tools/ncu_blocks.py rep [units] [min_share]
(units = what one 'unit of work' is, e.g. the number of chunks/tiles, to print executions per unit)"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.008
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[h]
ci, cs, cst = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
body = [(r[cs].strip(), int(r[ci]), int(r[cst]) if r[cst].isdigit() else 0) for r in rows[h + 1:] if len(r) > ci and r[ci].isdigit()]
tot = sum(b[1] for b in body)
tots = sum(b[2] for b in body) or 1
blocks, start = [], 0
for k in range(1, len(body) + 1):
    if k == len(body) or body[k][1] != body[start][1]:
        blocks.append((start, k, body[start][1]))
        start = k
print(f"# {len(body)} SASS instructions, {tot} warp-instructions ({tot / units:.1f} per unit), {tots} stall samples")
acc = 0.0
for s, e, n in blocks:
    share = (e - s) * n / tot
    st = sum(b[2] for b in body[s:e]) / tots
    if share < thr and st < thr:
        continue
    acc += share
    ops = [b[0].split()[1] if b[0].startswith("@") else b[0].split()[0] for b in body[s:e]]
    mem = [o for o in ops if o.startswith(("LDG", "STG", "LDS", "STS", "SHFL", "VOTE", "ATOM", "RED", "BRA", "CREDUX", "REDUX"))]
    print(f"[{s:4d},{e:4d}) {e - s:3d} instr x {n / units:7.2f}/unit = {100 * share:5.2f}% exec {100 * st:5.2f}% stall  {' '.join(mem[:12])}")
print(f"# listed blocks cover {100 * acc:.1f}% of executed instructions")
