#!/bin/bash
# Round 2, call T: stream priority of the detection path beside the build (high = default, normal), cooperative grid cap.
set -u
mkdir -p gpurun_out
for cfg in "-1 32" "0 32" "-1 8" "-1 148" "0 148"; do
  set -- $cfg
  GM_DET_PRIORITY=$1 GM_COOP_MAX_BLOCKS=$2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-iou --no-extras 2>> gpurun_out/r2t.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'prio': $1, 'coop': $2, 'ms_per_step': d['ms_per_step'], 'e2e_ms': d['e2e']['ms_per_step'], 'build_ms': d['roofline']['dtedge_build_ms'], 'merge_path_wall_ms': d['roofline']['merge_path_wall_ms']}))" >> gpurun_out/r2t.jsonl
done
cat gpurun_out/r2t.jsonl; tail -3 gpurun_out/r2t.err
