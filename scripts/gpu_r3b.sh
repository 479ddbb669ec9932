#!/bin/bash
# Launch list of one c5 device step of the FINAL binary (ncu --metrics gpu__time_duration.sum, all library kernels).
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-iou --no-extras"
$BENCH > gpurun_out/r3b_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'(::|^)k_' -c 3000 --csv --log-file gpurun_out/r3b_launches.csv $BENCH > gpurun_out/r3b_ncu_list.log 2>&1
tail -n 2 gpurun_out/r3b_ncu_list.log | cut -c 1-300; wc -l gpurun_out/r3b_launches.csv
