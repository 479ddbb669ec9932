#!/usr/bin/env python
"""DRAM bytes per launch of every kernel in an .ncu-rep -> JSON (read on the CPU box).
usage: ncu_traffic.py <rep> <out.json> "<source description>" [report.md ...]"""
import csv, io, json, subprocess, sys

rep, out, source = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
ALIAS = {"k_grad_fast": "k_grad"}          # bench.py's stage names
res, times = {}, {}
for r in data:
    name = r[ki].split("(")[0].split("::")[-1].split("<")[0]
    name = ALIAS.get(name, name)
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(m)
        tot += float(r[i].replace(",", "")) * SCALE[units[i]]
    res.setdefault(name, []).append(tot)
    i = hdr.index("gpu__time_duration.sum")
    times.setdefault(name, []).append(float(r[i].replace(",", "")) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}.get(units[i], 1.0))
doc = {"source": source, "reports": sys.argv[4:],
       "bytes_per_launch": {k: int(sum(v) / len(v)) for k, v in res.items()},
       "us_per_launch_under_ncu": {k: round(sum(v) / len(v), 1) for k, v in times.items()}}
with open(out, "w") as fh:
    json.dump(doc, fh, indent=1)
print(json.dumps(doc, indent=1))
