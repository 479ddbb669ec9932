#!/bin/bash
# Round 2, call X: bounded per-tile NMS (one CTA per tile, bit-mask sweep): parity against the engine and the oracle, the
# geometry / scale / seam / predictor suites, then the c5 step with and without it.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_geom.py tests/test_gpu_scale.py tests/test_gpu_seam.py tests/test_gpu_predictor.py tests/test_gpu_zz_adjacent.py tests/test_gpu_zz_yolo11.py -q 2>&1 | tail -12 > gpurun_out/r2x_pytest.log
cat gpurun_out/r2x_pytest.log
for f in 1 0; do
  GM_TILE_NMS_FAST=$f timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-iou --no-extras 2>> gpurun_out/r2x.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'fast': $f, 'ms_per_step': d['ms_per_step'], 'launches': d['gpu_launches'], 'tile_stage_ms': d['roofline']['tile_stage_ms'], 'merge_path_wall_ms': d['roofline']['merge_path_wall_ms'], 'build_ms': d['roofline']['dtedge_build_ms'], 'checksum': d['config']['merged_checksum'], 'survivors': d['config']['survivors_after_tile_nms'], 'adjacent': d['threshold_adjacent_pairs']}))" >> gpurun_out/r2x.jsonl
done
cat gpurun_out/r2x.jsonl; tail -3 gpurun_out/r2x.err
