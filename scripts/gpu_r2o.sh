#!/bin/bash
# Round 2, call O: full GPU suite on the one-kernel radix-sort offsets (3 launches per pass instead of 5) and the
# 40-register k_discover, then a quick c5 bench line (launch count, merge path time).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2o_pytest.log
cat gpurun_out/r2o_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2o_bench.json 2> gpurun_out/r2o.err; echo "rc=$?"
python -c "
import json
d = json.loads(open('gpurun_out/r2o_bench.json').read().strip().splitlines()[-1])
print(json.dumps({'ms_per_step': d['ms_per_step'], 'value': d['value'], 'e2e': d['e2e']['value'], 'launches': d['gpu_launches'], 'merge_path_wall_ms': d['roofline']['merge_path_wall_ms'], 'tile_stage_ms': d['roofline']['tile_stage_ms'], 'build_ms': d['roofline']['dtedge_build_ms'], 'checksum': d['config']['merged_checksum'], 'iou': d['iou']['gpairs_per_s'], 'iou_err': d['iou'].get('error_vs_float64')}))"
tail -3 gpurun_out/r2o.err
