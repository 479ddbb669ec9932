#!/bin/bash
# Round 2, call H: grad kernel restructure (parity + timing), IoU packed/scalar hybrids, FP32 pipe probe.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_tma.py tests/test_gpu_pixel.py tests/test_gpu_zy_otsu.py -q -x 2>&1 | tail -5 > gpurun_out/r2h_pytest.log
cat gpurun_out/r2h_pytest.log
for v in 1 0; do GM_GRAD_TMA=$v python scripts/probes/grad_leg.py >> gpurun_out/r2h_grad.jsonl 2>> gpurun_out/r2h.err; done
cat gpurun_out/r2h_grad.jsonl
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/f32x2_probe scripts/probes/f32x2_probe.cu && /tmp/f32x2_probe | tee gpurun_out/r2h_f32x2_probe.log
for v in q2fma q2fmaadd q2all; do GM_LIB_PATH=$PWD/oriented_object_detection_b200/lib/variants/$v.so GM_IOU_VARIANT=0 python scripts/probes/iou_leg.py >> gpurun_out/r2h_iou.jsonl 2>> gpurun_out/r2h.err; done
GM_IOU_VARIANT=0 python scripts/probes/iou_leg.py >> gpurun_out/r2h_iou.jsonl 2>> gpurun_out/r2h.err
GM_IOU_VARIANT=5 python scripts/probes/iou_leg.py >> gpurun_out/r2h_iou.jsonl 2>> gpurun_out/r2h.err
cat gpurun_out/r2h_iou.jsonl; tail -3 gpurun_out/r2h.err
