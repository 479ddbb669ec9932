#!/bin/bash
# Round 2, call F: packed IoU kernel + TMA gradient kernel: parity, timing A/B, ncu of both.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_tma.py tests/test_gpu_pixel.py tests/test_gpu_geom.py tests/test_gpu_zz_adjacent.py -q -x 2>&1 | tail -8 > gpurun_out/r2f_pytest.log
cat gpurun_out/r2f_pytest.log
cat > /tmp/iou_leg.py <<'PY'
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
print(json.dumps({"variant": os.environ.get("GM_IOU_VARIANT", "default"), "ms": ms, "gpairs": nb * nb / ms / 1e6, "checksum": float(rs.sum().item()), "ffma_peak": ops.ffma_peak(8192)}))
PY
cat > /tmp/grad_leg.py <<'PY'
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
PY
for v in 0 5; do GM_IOU_VARIANT=$v python /tmp/iou_leg.py >> gpurun_out/r2f_iou.jsonl 2>> gpurun_out/r2f.err; done
for v in 1 0; do GM_GRAD_TMA=$v python /tmp/grad_leg.py >> gpurun_out/r2f_grad.jsonl 2>> gpurun_out/r2f.err; done
cat gpurun_out/r2f_iou.jsonl gpurun_out/r2f_grad.jsonl; tail -3 gpurun_out/r2f.err
GM_IOU_VARIANT=0 python /tmp/iou_leg.py > gpurun_out/r2f_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_iou_matrix2' -c 2 -o gpurun_out/r2f_prof_iou2 python /tmp/iou_leg.py > gpurun_out/r2f_ncu.log 2>&1
python /tmp/grad_leg.py > gpurun_out/r2f_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_grad_fast' -s 2 -c 1 -o gpurun_out/r2f_prof_grad python /tmp/grad_leg.py > gpurun_out/r2f_ncu2.log 2>&1
tail -2 gpurun_out/r2f_ncu.log gpurun_out/r2f_ncu2.log | cut -c 1-300
