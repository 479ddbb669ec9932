"""Dense rotated-IoU throughput of the loaded library (GM_LIB_PATH / GM_IOU_VARIANT select the form): one JSON line."""
import json, os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from oriented_object_detection_b200 import ops, synth
dev = torch.device("cuda:0")
plan = ops.make_plan(8192, 8192, 416, 100, device=dev)
local, cls, conf, tid = synth.synthetic_tile_dets(plan, 59000, 15, seed=0, margin=20)
nb = 8192
b = local[:nb].astype(np.float64); b[:, 0::2] += plan.tiles["x0"][tid[:nb]][:, None]; b[:, 1::2] += plan.tiles["y0"][tid[:nb]][:, None]
bx = torch.from_numpy(b).to(dev); rs = torch.empty(nb, dtype=torch.float64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): ops.rotated_iou_matrix_sum(bx, bx, out=rs)
e0.record()
for _ in range(10): ops.rotated_iou_matrix_sum(bx, bx, out=rs)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(json.dumps({"lib": os.path.basename(os.environ.get("GM_LIB_PATH", "default")), "variant": os.environ.get("GM_IOU_VARIANT", "default"), "ms": ms,
                  "gpairs": nb * nb / ms / 1e6, "checksum": float(rs.sum().item()), "ffma_peak": ops.ffma_peak(8192)}))
