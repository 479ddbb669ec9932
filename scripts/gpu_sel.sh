#!/bin/bash
# selection variants: parity tests of the pixel path + stage timing; then ncu --set full of every DT-Edge kernel (one range)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pixel.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_pixel.log
cat gpurun_out/pytest_pixel.log
timeout 200 python scripts/sweep_select.py 2>&1 | head -3
export GM_DTEDGE_CHUNKS=1
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-iou"
ncu --set full --clock-control none --import-source on -k regex:"k_grad|k_select|k_edge|k_chamfer|k_tail" -s 12 -c 6 -o gpurun_out/prof_r1v4 $BENCH > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
