#!/usr/bin/env python
"""Per-instruction execution counts of the first kernel in an .ncu-rep (SASS view): prints the SASS with the
warp-level 'Instructions Executed' of each instruction as a share of the kernel total, collapsing cold
stretches, so loops that dominate the instruction stream stand out.

usage: tools/ncu_hot.py rep [min_share_pct]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[h]
ci, cs, cst = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
body = []
for r in rows[h + 1:]:
    if len(r) <= ci or not r[ci].isdigit():
        break
    body.append((r[cs].strip(), int(r[ci]), int(r[cst]) if r[cst].isdigit() else 0))
tot = sum(b[1] for b in body)
tots = sum(b[2] for b in body) or 1
print(f"# {len(body)} SASS instructions, {tot} warp-instructions executed, {tots} stall samples")
cold = 0
for k, (src, n, smp) in enumerate(body):
    share = 100.0 * n / tot
    if share >= thr:
        if cold:
            print(f"      ... {cold} colder instructions")
            cold = 0
        print(f"{k:5d} {share:5.2f}% exec {100.0 * smp / tots:5.2f}% stall  {src[:90]}")
    else:
        cold += 1
if cold:
    print(f"      ... {cold} colder instructions")
