"""Prints the per-stage times of the DT-Edge build (8192^2, 416/100) for the library / env in effect, and checks the
output against a checksum file shared by all variants of one gpurun call (first writer wins)."""
import os, sys, hashlib, torch
sys.path.insert(0, '.')
from oriented_object_detection_b200 import ops, synth
dev = torch.device('cuda:0')
H = W = 8192
plan = ops.make_plan(H, W, 416, 100, device=dev)
m = synth.synthetic_map(H, W, 1000, dev)
out = torch.empty(4 * plan.total_px, dtype=torch.uint8, device=dev)
for _ in range(3):
    ops.dtedge_build_timed(m, plan, out=out)
acc = {}
for _ in range(5):
    _, ms = ops.dtedge_build_timed(m, plan, out=out)
    for k, v in ms.items():
        acc[k] = acc.get(k, 0.0) + v / 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.dtedge_build(m, plan, out=out)
e1.record(); torch.cuda.synchronize()
digest = hashlib.sha1(out.cpu().numpy().tobytes()).hexdigest()[:12]
ref = "/tmp/stage_times.digest"
if not os.path.exists(ref):
    open(ref, "w").write(digest)
same = open(ref).read() == digest
print(f"{sys.argv[1] if len(sys.argv) > 1 else ''}: " + " ".join(f"{k} {v:.3f}" for k, v in acc.items()) +
      f" sum {sum(acc.values()):.3f} forked {e0.elapsed_time(e1) / 10:.3f} same={same}", flush=True)
