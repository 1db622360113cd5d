#!/usr/bin/env python
"""DRAM traffic per launch of one kernel from an `ncu --set full` report -> JSON (read by bench.py as
roofline.traffic).  usage: ncu_traffic.py report.ncu-rep kernel-substring out.json "<command that was profiled>" """
import csv, io, json, subprocess, sys
rep, sub, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
def col(name):
    return h.index(name)
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
recs = []
for r in rows[2:]:
    if sub not in r[col("Kernel Name")]:
        continue
    rd = float(r[col("dram__bytes_read.sum")]) * scale[units[col("dram__bytes_read.sum")]]
    wr = float(r[col("dram__bytes_write.sum")]) * scale[units[col("dram__bytes_write.sum")]]
    dur = float(r[col("gpu__time_duration.sum")])
    recs.append((rd, wr, dur, r[col("Kernel Name")]))
assert recs, "kernel not found in the report"
rd = sum(x[0] for x in recs) / len(recs)
wr = sum(x[1] for x in recs) / len(recs)
json.dump({"kernel": recs[0][3], "launches_captured": len(recs), "dram_bytes_read_per_launch": rd,
           "dram_bytes_write_per_launch": wr, "dram_bytes_per_launch": rd + wr,
           "duration_under_ncu": [x[2] for x in recs], "duration_unit": units[col("gpu__time_duration.sum")],
           "source": f"ncu --set full --clock-control none: dram__bytes_read.sum + dram__bytes_write.sum, {cmd} ({rep.split('/')[-1]})"},
          open(out, "w"), indent=1)
print(open(out).read())
