#!/bin/bash
# Round 2, call V: k_grad_fast walking several column blocks per CTA with the next TMA patch requested during the Scharr
# phase: parity with forced loop lengths, then c3 / c5 times against one block per CTA, at 7 and 6 CTAs per SM.
set -u
mkdir -p gpurun_out
for x in 13 3; do
  GM_GRAD_XLOOP=$x timeout 900 python -m pytest tests/test_gpu_pixel.py tests/test_gpu_zz_tma.py tests/test_gpu_zy_otsu.py tests/test_gpu_train.py -q 2>&1 | tail -3 >> gpurun_out/r2v_pytest.log
done
cat gpurun_out/r2v_pytest.log
for cfg in "default 1" "default 0" "default 13" "default 4" "grad6 1" "grad6 0" "grad6 13"; do
  set -- $cfg
  if [ $1 = default ]; then unset GM_LIB_PATH; else export GM_LIB_PATH=$PWD/oriented_object_detection_b200/lib/variants/$1.so; fi
  GM_GRAD_XLOOP=$2 python scripts/probes/grad_leg.py 2>> gpurun_out/r2v.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'lib': '$1', 'xloop': $2, 'grad': d['stages_ms']['grad'], 'build_ms': d['build_ms'], 'checksum': d['checksum']}))" >> gpurun_out/r2v_c3.jsonl
  GM_GRAD_XLOOP=$2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-iou --no-extras 2>> gpurun_out/r2v.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'lib': '$1', 'xloop': $2, 'ms_per_step': d['ms_per_step'], 'grad': d['roofline']['stages_ms']['grad'], 'build_ms': d['roofline']['dtedge_build_ms']}))" >> gpurun_out/r2v_c5.jsonl
done
unset GM_LIB_PATH
cat gpurun_out/r2v_c3.jsonl gpurun_out/r2v_c5.jsonl; tail -3 gpurun_out/r2v.err
