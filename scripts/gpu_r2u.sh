#!/bin/bash
# Round 2, call U: k_grad_fast at 7 / 8 CTAs per SM (32 registers, no spills) instead of 6 (39), with and without the
# maximum shared-memory carveout; c3 stage times and the c5 step.
set -u
mkdir -p gpurun_out
for cfg in "default -1" "grad7 -1" "grad8 -1" "grad8 100" "grad7 100"; do
  set -- $cfg
  if [ $1 = default ]; then unset GM_LIB_PATH; else export GM_LIB_PATH=$PWD/oriented_object_detection_b200/lib/variants/$1.so; fi
  GM_GRAD_CARVEOUT=$2 python scripts/probes/grad_leg.py 2>> gpurun_out/r2u.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'lib': '$1', 'carve': $2, 'grad': d['stages_ms']['grad'], 'build_ms': d['build_ms'], 'checksum': d['checksum']}))" >> gpurun_out/r2u_c3.jsonl
  GM_GRAD_CARVEOUT=$2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-iou --no-extras 2>> gpurun_out/r2u.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'lib': '$1', 'carve': $2, 'ms_per_step': d['ms_per_step'], 'grad': d['roofline']['stages_ms']['grad'], 'build_ms': d['roofline']['dtedge_build_ms']}))" >> gpurun_out/r2u_c5.jsonl
done
unset GM_LIB_PATH
cat gpurun_out/r2u_c3.jsonl gpurun_out/r2u_c5.jsonl; tail -3 gpurun_out/r2u.err
