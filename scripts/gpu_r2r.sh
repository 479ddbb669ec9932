#!/bin/bash
# Round 2, call R: occupancy caps of k_tail (80 -> 64 / 48 / 40 registers) and k_chamfer (98 -> 80 / 64) on c3 and the c5 step.
set -u
mkdir -p gpurun_out
for v in default tail4 tail5 tail6 cham12 cham16; do
  if [ $v = default ]; then unset GM_LIB_PATH; else export GM_LIB_PATH=$PWD/oriented_object_detection_b200/lib/variants/$v.so; fi
  python scripts/probes/grad_leg.py 2>> gpurun_out/r2r.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); d['lib'] = '$v'; print(json.dumps(d))" >> gpurun_out/r2r_c3.jsonl
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-iou --no-extras 2>> gpurun_out/r2r.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'lib': '$v', 'ms_per_step': d['ms_per_step'], 'stages_ms': d['roofline']['stages_ms'], 'build_ms': d['roofline']['dtedge_build_ms']}))" >> gpurun_out/r2r_c5.jsonl
done
unset GM_LIB_PATH
cat gpurun_out/r2r_c3.jsonl gpurun_out/r2r_c5.jsonl
tail -3 gpurun_out/r2r.err
