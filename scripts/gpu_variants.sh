#!/bin/bash
# bench under several values of a tuning env var. usage: gpu_variants.sh VAR v1 v2 ...
VAR=$1; shift
python -m pytest tests/test_gpu_pixel.py -m gpu -q -x 2>&1 | tail -3
for v in "$@"; do
  export $VAR=$v
  echo "== $VAR=$v"
  python -m pytest tests/test_gpu_pixel.py -m gpu -q -x -k "stage_by_stage or degenerate" 2>&1 | tail -1
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-iou 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['stages_ms'])
    else: print(l, end='')"
done
