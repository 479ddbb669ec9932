#!/bin/bash
# final lines of the round: bench (native + reference arm) and one ncu capture of the IoU kernel
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --set full --clock-control none --import-source on -k regex:'k_iou_matrix' -c 1 -o gpurun_out/prof_round_iou python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
cat gpurun_out/bench.json | cut -c1-200; tail -2 gpurun_out/bench.err
