#!/bin/bash
# Round 2, call I: full GPU suite + N=1 bench line + IoU occupancy variants + ncu launch list of the c5 step.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2i_pytest.log
cat gpurun_out/r2i_pytest.log
for v in iou8 iou9; do for m in 0 5; do GM_LIB_PATH=$PWD/oriented_object_detection_b200/lib/variants/$v.so GM_IOU_VARIANT=$m python scripts/probes/iou_leg.py >> gpurun_out/r2i_iou.jsonl 2>> gpurun_out/r2i.err; done; done
cat gpurun_out/r2i_iou.jsonl
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2i_bench.err
BENCH="python bench.py --steps 1 --warmup 3 --maps 2 --no-cpu-baseline --no-iou --no-extras"
$BENCH > gpurun_out/r2i_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'(::|^)k_' -c 3000 --csv --log-file gpurun_out/r2i_launches.csv $BENCH > gpurun_out/r2i_ncu_list.log 2>&1
tail -n 2 gpurun_out/r2i_ncu_list.log | cut -c 1-300
