#!/bin/bash
# One gpurun call: GPU tests, smoke, bench (native + reference arm), then the ncu launch list and one full
# capture of the pixel-path kernels (each ncu run directly after the same command exited 0 without ncu).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-iou"
$BENCH > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'(::|^)k_' -c 1400 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_list.log 2>&1
$BENCH > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_grad_fast|k_select_grad|k_edge_open|k_chamfer|k_select_dist|k_tail' -s 12 -c 6 -o gpurun_out/prof_round $BENCH > gpurun_out/ncu_full.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_iou_matrix' -c 2 -o gpurun_out/prof_round_iou python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench.err
