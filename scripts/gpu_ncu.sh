#!/bin/bash
# ncu --set full capture of the DT-Edge kernels (one launch each) after a plain run exited 0.
# usage: gpu_ncu.sh <regex> <out-name> [skip] [count]
set -u
mkdir -p gpurun_out
RE=${1:-k_grad}; OUT=${2:-prof}; SKIP=${3:-0}; CNT=${4:-6}
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-iou"
$BENCH > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -o gpurun_out/$OUT $BENCH > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
