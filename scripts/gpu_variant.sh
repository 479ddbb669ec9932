#!/bin/bash
# tests + bench under a tuning env var. usage: gpu_variant.sh VAR=VALUE
mkdir -p gpurun_out
export "$1"
python -m pytest tests/test_gpu_pixel.py -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-iou 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['stages_ms'])
    else: print(l, end='')"
