#!/bin/bash
# Round 2, call Q: selection CTAs of 512 / 256 threads with a proportionally smaller list (2-3 CTAs per SM instead of 1):
# parity of the pixel suite under each, stage times on the c5 step and on c3.
set -u
mkdir -p gpurun_out
for t in 512 256; do
  GM_SELECT_THREADS=$t timeout 900 python -m pytest tests/test_gpu_pixel.py tests/test_gpu_zz_tma.py -q 2>&1 | tail -3 >> gpurun_out/r2q_pytest.log
done
cat gpurun_out/r2q_pytest.log
for t in 1024 512 256; do
  GM_SELECT_THREADS=$t python scripts/probes/grad_leg.py 2>> gpurun_out/r2q.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); d['sel_threads'] = $t; print(json.dumps(d))" >> gpurun_out/r2q_c3.jsonl
  GM_SELECT_THREADS=$t timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-iou --no-extras 2>> gpurun_out/r2q.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'sel_threads': $t, 'ms_per_step': d['ms_per_step'], 'stages_ms': d['roofline']['stages_ms'], 'build_ms': d['roofline']['dtedge_build_ms'], 'merge_path_wall_ms': d['roofline']['merge_path_wall_ms']}))" >> gpurun_out/r2q_c5.jsonl
done
cat gpurun_out/r2q_c3.jsonl gpurun_out/r2q_c5.jsonl
tail -3 gpurun_out/r2q.err
