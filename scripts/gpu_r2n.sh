#!/bin/bash
# Round 2, call N: register cap of k_discover beyond 8 CTAs / SM, then the full bench line and the reference arm.
set -u
mkdir -p gpurun_out
for v in default disc10n disc12n disc16n; do
  if [ $v = default ]; then unset GM_LIB_PATH; else export GM_LIB_PATH=$PWD/oriented_object_detection_b200/lib/variants/$v.so; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-iou --no-extras 2>> gpurun_out/r2n.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'lib': '$v', 'ms_per_step': d['ms_per_step'], 'merge_path_wall_ms': d['roofline']['merge_path_wall_ms'], 'tile_stage_ms': d['roofline']['tile_stage_ms'], 'build_ms': d['roofline']['dtedge_build_ms'], 'checksum': d['config']['merged_checksum']}))" >> gpurun_out/r2n_discover.jsonl
done
unset GM_LIB_PATH
cat gpurun_out/r2n_discover.jsonl
timeout 900 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2n_bench.err
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2n_bench_ref.json 2> gpurun_out/r2n_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/r2n_bench_ref.json | cut -c 1-600
tail -5 gpurun_out/r2n.err
