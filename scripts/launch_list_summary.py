#!/usr/bin/env python
"""One step of the ncu launch list (`--metrics gpu__time_duration.sum`, all kernels) as a table: the launches between
two consecutive k_grad_fast launches, grouped by kernel name, with each group's share of the step.
usage: launch_list_summary.py <launches.csv> <which step (0-based k_grad_fast index)> <out.md> <out.csv>"""
import csv, sys
src, which, out_md, out_csv = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.DictReader(lines))
def short(n):
    n = n.split("(")[0]
    n = n.replace("void ", "").replace("<unnamed>::", "")
    return n[-70:]
grads = [i for i, r in enumerate(rows) if "k_grad_fast" in r["Kernel Name"]]
a, b = grads[which], grads[which + 1]
step = rows[a:b]
def ns(r):
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    return v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "ns ": 1.0}.get(u, 1.0)
tot = sum(ns(r) for r in step)
agg = {}
for r in step:
    k = short(r["Kernel Name"])
    c, t = agg.get(k, (0, 0.0))
    agg[k] = (c + 1, t + ns(r))
own = sum(c for k, (c, t) in agg.items() if k.startswith("k_") or "::k_" in k or "gradfast" in k)
with open(out_md, "w") as fh:
    fh.write(f"# ncu launch list of one device step ({len(step)} launches, {own} of them library kernels; serialised, cold-cache times: {tot / 1e6:.3f} ms)\n\n")
    fh.write("Source: `ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py --steps 2 --warmup 3 --no-cpu-baseline`; "
             f"launches {a}..{b - 1} of the list (from one k_grad_fast to the next).  Under ncu the two streams of a step are serialised and every "
             "kernel starts cold, so the absolute times are larger than in the timed run; the SHARE per kernel is what carries over.\n\n")
    fh.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write(f"| `{k}` | {c} | {t / 1e3:.1f} | {100 * t / tot:.1f} % |\n")
with open(out_csv, "w") as fh:
    w = csv.writer(fh)
    w.writerow(["index", "kernel", "stream", "block", "grid", "ns"])
    for i, r in enumerate(step):
        w.writerow([a + i, short(r["Kernel Name"]), r["Stream"], r["Block Size"], r["Grid Size"], int(ns(r))])
print(open(out_md).read()[:3000])
