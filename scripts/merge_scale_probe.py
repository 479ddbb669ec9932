"""Where the cross-band merge spends its time at N ranks, emulated on ONE GPU: the all_gather result is assembled from N
synthetic bands, then the steps every rank runs on world*capacity rows are timed (eager, synchronous) and the whole
device part is replayed as a CUDA graph."""
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from oriented_object_detection_b200 import ops, synth, sharding
dev = torch.device('cuda:0')
W = 8192
def t(a): return torch.from_numpy(a).to(dev)
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / reps, r
for world in (1, 4, 8):
    H = 8192 * world
    plan = ops.make_plan(H, W, 416, 100, device=dev)
    local, cls, conf, tid = synth.synthetic_tile_dets(plan, 59000 * world, 15, seed=0, margin=20)
    pp = ops.tile_postprocess(t(local), t(cls), t(conf), t(tid), plan, 20, 1, 0.4, max_class=14)
    n = pp["conf"].shape[0]
    cap = (n + world - 1) // world + 1024
    total = world * cap
    # the gathered byte matrix: rows in band order, padded per band like the real exchange (approximation: equal split)
    keys = ["boxes", "cls", "conf", "angle"]
    fields = {}
    for k in keys:
        v = pp[k]
        blank = float("nan") if k == "boxes" else (-1 if k == "cls" else (float("-inf") if k == "conf" else 0.0))
        full = torch.full((total,) + tuple(v.shape[1:]), blank, dtype=v.dtype, device=dev)
        full[:n] = v
        fields[k] = full
    cols = [fields[k].reshape(total, -1).view(torch.uint8).reshape(total, -1) for k in keys]
    widths = [c.shape[1] for c in cols]
    recv = torch.cat(cols, dim=1).contiguous()
    def unpack():
        out, off = {}, 0
        for k, wd in zip(keys, widths):
            tail = tuple(fields[k].shape[1:])
            out[k] = recv[:, off:off + wd].contiguous().view(fields[k].dtype).reshape((total,) + tail)
            off += wd
        c = out["cls"]
        mine = (c >= 0) & ((c % world) == 0)
        out["mine"] = mine
        out["cls_mine"] = torch.where(mine, c, torch.full_like(c, -1))
        return out
    ms_unpack, u = timeit(unpack)
    ms_nms, r = timeit(lambda: ops.nms_global(u["boxes"], u["cls_mine"], u["conf"], 0.4, max_class=14, sync=False))
    order, keep = r[0], r[1]
    def extract():
        k8 = (keep.to(torch.uint8) * u["mine"].to(torch.uint8)).contiguous()
        o = order.to(torch.int64)
        flags = k8[o].to(torch.bool)
        pos = torch.cumsum(flags.to(torch.int64), 0) - 1
        slot = torch.where(flags, pos, torch.full_like(pos, total))
        kp = torch.zeros(total + 1, dtype=torch.int64, device=dev)
        kp.scatter_(0, slot, o)
        idx = kp[:total]
        return {k: u[k][idx] for k in keys}
    ms_extract, _ = timeit(extract)
    def whole():
        uu = unpack()
        rr = ops.nms_global(uu["boxes"], uu["cls_mine"], uu["conf"], 0.4, max_class=14, sync=False)
        return rr
    call = sharding.CapturedCall(lambda: {"x": extract()["conf"], "y": whole()[1]})
    ms_graph, _ = timeit(lambda: call())
    print(f"world {world}: {total} rows | eager+sync: unpack/mask {ms_unpack:.3f}  nms {ms_nms:.3f}  extract {ms_extract:.3f} ms | "
          f"graph replay of unpack+nms+extract {ms_graph:.3f} ms (captured={call.captured})", flush=True)
