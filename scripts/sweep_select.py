"""Stage times of the DT-Edge build for env-selected variants (one GPU); output checked against the first."""
import os, sys, ctypes as C, torch
sys.path.insert(0, '.')
import __graft_entry__ as entry
entry.build()
from oriented_object_detection_b200 import ops, synth, _lib
dev = torch.device('cuda:0')
H = W = 8192
plan = ops.make_plan(H, W, 416, 100, device=dev)
m = synth.synthetic_map(H, W, 1000, dev)
out = torch.empty(4 * plan.total_px, dtype=torch.uint8, device=dev)
ref = None
configs = [dict(kv.split("=") for kv in a.split(",")) if a != "-" else {} for a in sys.argv[1:]] or [{}]
keys = sorted({k for c in configs for k in c})
for cfg in configs:
    for k in keys:
        os.environ.pop(k, None)
    os.environ.update(cfg)
    for _ in range(2):
        ops.dtedge_build_timed(m, plan, out=out)
    acc = {}
    for _ in range(5):
        _, ms = ops.dtedge_build_timed(m, plan, out=out)
        for k, v in ms.items():
            acc[k] = acc.get(k, 0.0) + v / 5
    if ref is None:
        ref = out.clone()
    print(f"{cfg}: " + " ".join(f"{k} {v:.3f}" for k, v in acc.items()) + f" total {sum(acc.values()):.3f} equal={bool(torch.equal(out, ref))}", flush=True)
