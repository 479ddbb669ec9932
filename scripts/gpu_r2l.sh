#!/bin/bash
# Round 2, call L: tile-major grid of k_grad_fast (parity + time + DRAM traffic), float64 pair entry point,
# register-cap variants of k_discover (detection path wall time) and of the packed IoU kernel.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2l_pytest.log
cat gpurun_out/r2l_pytest.log
for t in 1 0; do GM_GRAD_TMA=$t python scripts/probes/grad_leg.py >> gpurun_out/r2l_grad.jsonl 2>> gpurun_out/r2l.err; done
cat gpurun_out/r2l_grad.jsonl
python scripts/probes/iou_leg.py >> gpurun_out/r2l_iou.jsonl 2>> gpurun_out/r2l.err
for v in ioumb6 ioumb7; do GM_LIB_PATH=$PWD/oriented_object_detection_b200/lib/variants/$v.so python scripts/probes/iou_leg.py >> gpurun_out/r2l_iou.jsonl 2>> gpurun_out/r2l.err; done
cat gpurun_out/r2l_iou.jsonl
for v in default disc3 disc4 disc5; do
  if [ $v = default ]; then unset GM_LIB_PATH; else export GM_LIB_PATH=$PWD/oriented_object_detection_b200/lib/variants/$v.so; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-iou --no-extras 2>> gpurun_out/r2l.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'lib': '$v', 'ms_per_step': d['ms_per_step'], 'merge_path_wall_ms': d['roofline']['merge_path_wall_ms'], 'tile_stage_ms': d['roofline']['tile_stage_ms'], 'build_ms': d['roofline']['dtedge_build_ms'], 'grad': d['roofline']['stages_ms']['grad'], 'checksum': d['config']['merged_checksum']}))" >> gpurun_out/r2l_discover.jsonl
done
unset GM_LIB_PATH
cat gpurun_out/r2l_discover.jsonl
python scripts/probes/grad_leg.py > gpurun_out/r2l_plain_grad.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_grad_fast' -s 2 -c 1 -o gpurun_out/r2l_prof_grad python scripts/probes/grad_leg.py > gpurun_out/r2l_ncu_grad.log 2>&1
tail -n 2 gpurun_out/r2l_ncu_grad.log | cut -c 1-200
tail -5 gpurun_out/r2l.err
