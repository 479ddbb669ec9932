#!/bin/bash
# Round 2, call M: full GPU suite (noinline float64 paths in k_discover, staged k_prepare stores, tile-major gradient
# grid, float64 pair entry point) + register-cap variants of k_discover on the c5 step.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2m_pytest.log
cat gpurun_out/r2m_pytest.log
for v in default disc5n disc6n disc8n; do
  if [ $v = default ]; then unset GM_LIB_PATH; else export GM_LIB_PATH=$PWD/oriented_object_detection_b200/lib/variants/$v.so; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-iou --no-extras 2>> gpurun_out/r2m.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'lib': '$v', 'ms_per_step': d['ms_per_step'], 'merge_path_wall_ms': d['roofline']['merge_path_wall_ms'], 'tile_stage_ms': d['roofline']['tile_stage_ms'], 'build_ms': d['roofline']['dtedge_build_ms'], 'grad': d['roofline']['stages_ms']['grad'], 'checksum': d['config']['merged_checksum']}))" >> gpurun_out/r2m_discover.jsonl
done
unset GM_LIB_PATH
cat gpurun_out/r2m_discover.jsonl
tail -5 gpurun_out/r2m.err
