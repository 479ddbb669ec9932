#!/bin/bash
set -u
mkdir -p gpurun_out
for b in 300 0; do GM_LEG_MAX_DET=$b python scripts/probes/tile_stage_leg.py >> gpurun_out/r2y.jsonl 2>> gpurun_out/r2y.err; done
cat gpurun_out/r2y.jsonl
GM_LEG_MAX_DET=300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'(::|^)k_' -c 400 --csv --log-file gpurun_out/r2y_launches.csv python scripts/probes/tile_stage_leg.py > gpurun_out/r2y_ncu.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2y_launches.csv')) if len(r) > 10 and r[0].isdigit()]
hdr = None
for r in csv.reader(open('gpurun_out/r2y_launches.csv')):
    if 'Kernel Name' in r: hdr = r; break
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
seq = [(r[ki].split('(')[0].split('::')[-1], float(r[vi].replace(',', ''))) for r in rows]
# last call = last occurrence of k_tile_remap onwards
last = max(i for i, (k, _) in enumerate(seq) if k.startswith('k_tile_remap'))
tot = 0
for k, v in seq[last:]:
    print(f"{k:28s} {v/1000:9.1f} us"); tot += v
print('total', tot / 1000, 'us')
PY
tail -3 gpurun_out/r2y.err
