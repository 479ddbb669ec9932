"""Sweep of the chunk x stream split of gm_dtedge_build_u8 (GM_DTEDGE_CHUNKS / GM_DTEDGE_STREAMS / GM_DTEDGE_SKEW) on one
GPU: 8192^2 synthetic map, 416/100 tiling; output checked equal to the single-range build."""
import os, sys, torch
sys.path.insert(0, '.')
import __graft_entry__ as entry
entry.build()
from oriented_object_detection_b200 import ops, synth
dev = torch.device('cuda:0')
H = W = 8192
plan = ops.make_plan(H, W, 416, 100, device=dev)
m = synth.synthetic_map(H, W, 1000, dev)
out = torch.empty(4 * plan.total_px, dtype=torch.uint8, device=dev)
os.environ["GM_DTEDGE_CHUNKS"] = "1"; os.environ["GM_DTEDGE_STREAMS"] = "1"
ref = ops.dtedge_build(m, plan).clone()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for skew in (0, 1):
    for chunks, lanes in [(1, 1), (2, 2), (3, 2), (3, 3), (4, 2), (4, 3), (4, 4), (6, 2), (6, 3), (6, 6), (8, 2), (8, 3), (8, 4), (13, 2), (13, 3), (13, 4)]:
        if skew == 0 and chunks > 4:
            continue
        os.environ["GM_DTEDGE_CHUNKS"] = str(chunks); os.environ["GM_DTEDGE_STREAMS"] = str(lanes); os.environ["GM_DTEDGE_SKEW"] = str(skew)
        out.zero_()
        for _ in range(3):
            ops.dtedge_build(m, plan, out=out)
        torch.cuda.synchronize()
        ok = bool(torch.equal(out, ref))
        e0.record()
        for _ in range(10):
            ops.dtedge_build(m, plan, out=out)
        e1.record(); torch.cuda.synchronize()
        print(f"skew {skew} chunks {chunks:2d} streams {lanes}: {e0.elapsed_time(e1) / 10:.3f} ms  equal={ok}", flush=True)
