"""Stage times of the DT-Edge build on BASELINE config 3 (8192^2, 416/100): one JSON line (GM_GRAD_TMA=0/1 selects the load path)."""
import json, os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from oriented_object_detection_b200 import ops, synth
dev = torch.device("cuda:0")
plan = ops.make_plan(8192, 8192, 416, 100, device=dev)
m = synth.synthetic_map(8192, 8192, 1000, dev)
out = torch.empty(4 * plan.total_px, dtype=torch.uint8, device=dev)
acc = {}
for _ in range(6):
    _, ms = ops.dtedge_build_timed(m, plan, out=out)
    for k, v in ms.items(): acc[k] = min(acc.get(k, 1e9), v)
print(json.dumps({"tma": os.environ.get("GM_GRAD_TMA", "1"), "stages_ms": acc, "build_ms": sum(acc.values()), "checksum": int(out[::4097].to(torch.int64).sum().item())}))
