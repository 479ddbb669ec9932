"""Stage times of the DT-Edge build for selection / chamfer variants + sampled-path stats (one GPU)."""
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
configs = [{"GM_SELECT_SAMPLED": "0"}, {"GM_SELECT_SAMPLED": "1"}, {"GM_SELECT_SAMPLED": "1", "GM_SELECT_THREADS": "512"}]
configs += [{"GM_CHAMFER_VARIANT": str(v)} for v in (1, 3, 4, 5)]
for cfg in configs:
    for k in ("GM_SELECT_SAMPLED", "GM_SELECT_THREADS", "GM_CHAMFER_VARIANT"):
        os.environ.pop(k, None)
    os.environ.update(cfg)
    for _ in range(2):
        ops.dtedge_build_timed(m, plan, out=out)
    cnt = (C.c_uint32 * 8)()
    _lib.lib.gm_dtedge_select_stats(cnt, 1)
    acc = {}
    for _ in range(5):
        _, ms = ops.dtedge_build_timed(m, plan, out=out)
        for k, v in ms.items():
            acc[k] = acc.get(k, 0.0) + v / 5
    _lib.lib.gm_dtedge_select_stats(cnt, 1)
    if ref is None:
        ref = out.clone()
    print(f"{cfg}: " + " ".join(f"{k} {v:.3f}" for k, v in acc.items()) + f" total {sum(acc.values()):.3f} stats/5 runs {list(cnt)} "
          f"equal={bool(torch.equal(out, ref))}", flush=True)
