#!/bin/bash
# Round 2, call J: full GPU suite on the new IoU forms (fp32 reference points, min/max slab cuts, prepared records),
# IoU variants, gradient leg (unconditional horizontal-pass stores), chamfer variants in the c5 (21,632-tile) regime.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2j_pytest.log
cat gpurun_out/r2j_pytest.log
for m in 0 3 5; do GM_IOU_VARIANT=$m python scripts/probes/iou_leg.py >> gpurun_out/r2j_iou.jsonl 2>> gpurun_out/r2j.err; done
cat gpurun_out/r2j_iou.jsonl
for t in 1 0; do GM_GRAD_TMA=$t python scripts/probes/grad_leg.py >> gpurun_out/r2j_grad.jsonl 2>> gpurun_out/r2j.err; done
cat gpurun_out/r2j_grad.jsonl
for v in 0 1 2 15; do
  GM_CHAMFER_VARIANT=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-iou --no-extras 2>> gpurun_out/r2j.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'chamfer_variant': $v, 'ms_per_step': d['ms_per_step'], 'stages_ms': d['roofline']['stages_ms'], 'build_ms': d['roofline']['dtedge_build_ms']}))" >> gpurun_out/r2j_chamfer.jsonl
done
cat gpurun_out/r2j_chamfer.jsonl
tail -5 gpurun_out/r2j.err
