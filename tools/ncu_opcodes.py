#!/usr/bin/env python
"""Per-opcode and per-region instruction counts of one kernel from an .ncu-rep (SASS source page).
usage: ncu_opcodes.py report.ncu-rep [units]   (units: divide the counts by this, e.g. patches per launch)"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = raw.splitlines()
hdr = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[hdr:]))))
h = rows[0]
ci, cs, cw = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
ops, samp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[1:]:
    if len(r) <= ci or not r[ci].isdigit():
        continue
    src = r[cs].split()
    op = src[1] if src and src[0].startswith("@") else (src[0] if src else "?")
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LD", "ST", "MUFU", "SHFL")) and "." in op else "")
    n = int(r[ci]); ops[op] += n; tot += n
    samp[op] += int(r[cw]) if r[cw].isdigit() else 0
print(f"total {tot / units:.1f} instructions per unit; static SASS instructions {len(rows) - 1}")
ts = sum(samp.values())
for op, n in ops.most_common(28):
    print(f"  {op:14s} {n / units:9.1f}  ({100.0 * n / tot:4.1f} %)   stall samples {100.0 * samp[op] / max(ts, 1):4.1f} %")
