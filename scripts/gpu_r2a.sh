#!/bin/bash
# Round 2, call A: GPU tests after the hygiene changes, bench (fixed FFMA probe), ncu --set full of the shipped dense IoU kernel.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2a_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_iou_matrix' -c 2 -o gpurun_out/r2a_prof_iou python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu.log 2>&1
tail -4 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
