#!/bin/bash
# Round 2, call K: ncu --set full evidence of what ships: dense IoU (packed, prepared records), the six DT-Edge kernels
# on c3 (DRAM traffic per launch -> profiles/r02_traffic.json), the detection-path kernels on a 2-map c5 step.
set -u
mkdir -p gpurun_out
python scripts/probes/iou_leg.py > gpurun_out/r2k_plain_iou.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_iou_matrix2' -s 2 -c 1 -o gpurun_out/r2k_prof_iou python scripts/probes/iou_leg.py > gpurun_out/r2k_ncu_iou.log 2>&1
python scripts/probes/grad_leg.py > gpurun_out/r2k_plain_grad.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_grad_fast|k_select_grad|k_edge_open|k_chamfer|k_select_dist|k_tail' -s 12 -c 6 -o gpurun_out/r2k_prof_dtedge python scripts/probes/grad_leg.py > gpurun_out/r2k_ncu_dtedge.log 2>&1
BENCH="python bench.py --steps 1 --warmup 3 --maps 2 --no-cpu-baseline --no-iou --no-extras"
$BENCH > gpurun_out/r2k_plain_bench.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_discover|k_nms_fixpoint|k_prepare|k_rs_scatter|k_rs_hist' -s 20 -c 12 -o gpurun_out/r2k_prof_merge $BENCH > gpurun_out/r2k_ncu_merge.log 2>&1
tail -n 2 gpurun_out/r2k_ncu_iou.log gpurun_out/r2k_ncu_dtedge.log gpurun_out/r2k_ncu_merge.log | cut -c 1-300
ls -la gpurun_out/r2k_*
