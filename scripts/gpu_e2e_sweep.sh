#!/bin/bash
python -m pytest tests/test_gpu_pixel.py -m gpu -q -x -k "streamed" 2>&1 | tail -2
for c in 3 4 6 8 13; do
  echo "== chunks=$c"
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-iou --chunks $c 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('dev', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
    else: print(l, end='')"
done
