#!/usr/bin/env python
"""Per-phase (between CTA barriers) instruction and stall-sample shares of one kernel from an
.ncu-rep source page.  usage: ncu_phases.py <rep> <kernel regex> <units for per-unit count>"""
import csv, io, subprocess, sys
rep, rx, units = sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; seg = 0; tot = {}; smp = {}; first = True
for r in rows:
    if len(r) > 5 and r[0] == "Address":
        if hdr is not None: break          # first launch only
        hdr = r; iS = r.index("Source"); iI = r.index("Instructions Executed"); iM = r.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    tot[seg] = tot.get(seg, 0) + int(r[iI]); smp[seg] = smp.get(seg, 0) + int(r[iM])
    if "BAR.SYNC" in r[iS]: seg += 1
T = sum(tot.values()); S = max(1, sum(smp.values()))
for k in tot:
    print(f"phase {k}: warp-instr {tot[k]:>12} ({tot[k]/T:5.1%})  samples {smp[k]/S:5.1%}  thread-instr/unit {tot[k]*32/units:7.1f}")
print(f"total warp-instr {T}, thread-instr/unit {T*32/units:.1f}")
