"""Times the NMS engine on N x capacity padded rows with (N-1)/N of the classes masked (what every rank runs
in the padded cross-band merge at N GPUs), on one GPU."""
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from oriented_object_detection_b200 import ops, synth
dev = torch.device('cuda:0')
W = 8192
for world in (1, 2, 4, 8):
    H = 8192 * world
    plan = ops.make_plan(H, W, 416, 100, device=dev)
    local, cls, conf, tid = synth.synthetic_tile_dets(plan, 59000 * world, 15, seed=0, margin=20)
    t = lambda a: torch.from_numpy(a).to(dev)
    pp = ops.tile_postprocess(t(local), t(cls), t(conf), t(tid), plan, 20, 1, 0.4, max_class=14)
    n = pp["conf"].shape[0]
    cls_m = torch.where((pp["cls"] % world) == 0, pp["cls"], torch.full_like(pp["cls"], -1))
    for name, c in (("all classes", pp["cls"]), ("1/N classes", cls_m)):
        for _ in range(3):
            ops.nms_global(pp["boxes"], c, pp["conf"], 0.4, max_class=14, sync=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.nms_global(pp["boxes"], c, pp["conf"], 0.4, max_class=14, sync=False)
        e1.record(); torch.cuda.synchronize()
        print(f"world {world}: {n} rows, {name}: {e0.elapsed_time(e1)/5:.3f} ms")
