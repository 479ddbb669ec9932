#!/bin/bash
# Round 2, last evidence call: ncu --set full of the shipped kernels after the occupancy caps (six DT-Edge kernels on c3),
# the two fused small-tile kernels on the 128/30 plan, the final k_tile_nms, the final dense IoU kernel.
set -u
mkdir -p gpurun_out
python scripts/probes/grad_leg.py > gpurun_out/r2zz_plain_grad.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_grad_fast|k_select_grad|k_edge_open|k_chamfer|k_select_dist|k_tail' -s 12 -c 6 -o gpurun_out/r2zz_prof_dtedge python scripts/probes/grad_leg.py > gpurun_out/r2zz_ncu_dtedge.log 2>&1
cat > /tmp/leg128.py <<'PY'
import json, os, sys, torch
sys.path.insert(0, os.getcwd())
from oriented_object_detection_b200 import ops, synth
dev = torch.device("cuda:0")
plan = ops.make_plan(8192, 8192, 128, 30, device=dev)
m = synth.synthetic_map(8192, 8192, 1000, dev)
out = torch.empty(4 * plan.total_px, dtype=torch.uint8, device=dev)
for _ in range(3):
    _, ms = ops.dtedge_build_timed(m, plan, out=out)
print(json.dumps(ms))
PY
python /tmp/leg128.py > gpurun_out/r2zz_plain_128.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_select_edge_small|k_select_tail_small' -s 2 -c 2 -o gpurun_out/r2zz_prof_small python /tmp/leg128.py > gpurun_out/r2zz_ncu_small.log 2>&1
GM_LEG_MAX_DET=300 ncu --set full --clock-control none --import-source on -k regex:k_tile_nms -s 2 -c 1 -o gpurun_out/r2zz_prof_tilenms python scripts/probes/tile_stage_leg.py > gpurun_out/r2zz_ncu_tilenms.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_iou_matrix2|k_iou_prepare' -s 4 -c 2 -o gpurun_out/r2zz_prof_iou python scripts/probes/iou_leg.py > gpurun_out/r2zz_ncu_iou.log 2>&1
ls -la gpurun_out/r2zz_*.ncu-rep
