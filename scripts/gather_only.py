import sys, torch
sys.path.insert(0, '.')
from oriented_object_detection_b200 import ops, synth
dev = torch.device('cuda:0')
H = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m = synth.synthetic_map(H, H, 1000, dev)
plan = ops.make_plan(H, H, 416, 100, device=dev)
out = torch.empty(3 * plan.total_px, dtype=torch.uint8, device=dev)
for _ in range(4):
    ops.tile_gather(m, plan, out=out)
torch.cuda.synchronize()
print("ok")
