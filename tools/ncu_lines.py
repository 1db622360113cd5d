#!/usr/bin/env python
"""Attribute the executed instructions (and stall samples) of one kernel to SOURCE LINES: ncu's SASS page gives the
per-instruction counts, nvdisasm -g on the object file gives the line of every instruction (same order).
usage: ncu_lines.py report.ncu-rep object.o mangled-substring [units] [top]"""
import csv, subprocess, sys, collections, io, re, tempfile, os, glob
rep, obj, sub = sys.argv[1], sys.argv[2], sys.argv[3]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = raw.splitlines()
hdr = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[hdr:]))))
h = rows[0]
ci, cs, cw = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
inst = [(r[cs], int(r[ci]), int(r[cw]) if r[cw].isdigit() else 0) for r in rows[1:] if len(r) > ci and r[ci].isdigit()]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and sub in l)
cur, per_inst = None, []
for l in dis[start + 1:]:
    if l.startswith("//----") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        per_inst.append(cur)
n = min(len(per_inst), len(inst))
if len(per_inst) != len(inst):
    print(f"warning: {len(per_inst)} disassembled vs {len(inst)} profiled instructions")
agg, smp = collections.Counter(), collections.Counter()
for k in range(n):
    agg[per_inst[k]] += inst[k][1]
    smp[per_inst[k]] += inst[k][2]
tot, ts = sum(agg.values()), sum(smp.values())
print(f"total {tot / units:.1f} instructions per unit")
srcs = {}
for (f, ln), c in agg.most_common(top):
    if f not in srcs:
        cand = glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "geostatsmodels.jl_b200", "csrc", f))
        srcs[f] = open(cand[0]).read().splitlines() if cand else []
    text = srcs[f][ln - 1].strip()[:90] if ln - 1 < len(srcs[f]) else ""
    print(f"{c / units:8.1f} {100.0 * c / tot:5.1f}%  stall {100.0 * smp[(f, ln)] / max(ts, 1):5.1f}%  {f}:{ln}  {text}")
