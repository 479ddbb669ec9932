#!/usr/bin/env python
"""Hot SASS regions of one kernel of an .ncu-rep (source page): usage ncu_hot.py <rep> <kernel-name> [chunk]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[2:]:
    if len(r) < 10:
        break                       # second launch of the same kernel
    if r[iex].isdigit():
        data.append(r)
tot = sum(int(r[iex]) for r in data); ts = sum(int(r[ismp]) for r in data)
print(f"{kern}: {tot} warp instructions, {len(data)} SASS lines, {ts} samples")
for i in range(0, len(data), chunk):
    c = sum(int(r[iex]) for r in data[i:i + chunk]); sm = sum(int(r[ismp]) for r in data[i:i + chunk])
    if c * 100 > tot or sm * 100 > ts:
        ops = {}
        for r in data[i:i + chunk]:
            op = r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]
            ops[op] = ops.get(op, 0) + int(r[iex])
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
        print(f"  sass {i:5d}: {100 * c / tot:5.1f}% instr {100 * sm / max(ts, 1):5.1f}% samples  " + " ".join(f"{k}:{v * 100 // max(c, 1)}%" for k, v in top))
