"""Per-tile stage (remap + border filter + per-tile NMS) on the detections of `GM_LEG_MAPS` 16384^2 maps: one JSON line
(GM_LEG_MAX_DET = 0 takes the general engine, 300 the bounded one-CTA-per-tile form)."""
import json, os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from oriented_object_detection_b200 import ops, synth
dev = torch.device("cuda:0")
M = int(os.environ.get("GM_LEG_MAPS", "2"))
bound = int(os.environ.get("GM_LEG_MAX_DET", "300"))
full = ops.make_plan(16384, 16384, 416, 100)
parts = []
for m in range(M):
    local, cls, conf, tid = synth.synthetic_tile_dets(full, 236000, 15, seed=m, margin=20)
    parts.append((local, cls, conf, (tid.astype(np.int64) + m * full.n).astype(np.int32)))
dets = [torch.from_numpy(np.ascontiguousarray(np.concatenate([p[k] for p in parts]))).to(dev) for k in range(4)]
rep = lambda a: np.tile(np.asarray(a), M)
plan = ops.plan_from_arrays(16384, 16384, rep(full.tiles["y0"]), rep(full.tiles["x0"]), rep(full.tiles["h"]), rep(full.tiles["w"]), device=dev,
                            tile_size=416, overlap=100)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(4):
    torch.cuda.synchronize(); e0.record()
    pp = ops.tile_postprocess(dets[0], dets[1], dets[2], dets[3], plan, 20, 1, 0.4, max_class=14, sync=False, max_per_tile=bound)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"maps": M, "boxes": int(dets[2].numel()), "max_per_tile": bound, "ms": best, "survivors": int(pp["count"].item()),
                  "max_count": int(np.bincount(dets[3].cpu().numpy()).max())}))
