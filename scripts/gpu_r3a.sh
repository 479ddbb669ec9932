#!/bin/bash
# Selection kernels with more private list slots per thread (one CTA per SM fits regardless: registers) and wider
# brackets for the distance percentiles: fallback counts and stage times on c3 and the c5 step.
set -u
mkdir -p gpurun_out
cat > /tmp/sel_stats.py <<'PY'
import json, os, sys, ctypes as C, torch
sys.path.insert(0, os.getcwd())
from oriented_object_detection_b200 import ops, synth, _lib
dev = torch.device("cuda:0")
plan = ops.make_plan(8192, 8192, 416, 100, device=dev)
m = synth.synthetic_map(8192, 8192, 1000, dev)
out = torch.empty(4 * plan.total_px, dtype=torch.uint8, device=dev)
cnt = (C.c_uint32 * 8)()
_lib.lib.gm_dtedge_select_stats(cnt, 1)
acc = {}
for _ in range(5):
    _, ms = ops.dtedge_build_timed(m, plan, out=out)
    for k, v in ms.items(): acc[k] = min(acc.get(k, 1e9), v)
_lib.lib.gm_dtedge_select_stats(cnt, 0)
print(json.dumps({"lib": os.path.basename(os.environ.get("GM_LIB_PATH", "default")), "select_grad": acc["select_grad"], "select_dist": acc["select_dist"], "build_ms": sum(acc.values()),
                  "stats_per_5_builds": list(cnt), "checksum": int(out[::4097].to(torch.int64).sum().item())}))
PY
for v in default sel44 sel44s16 sel36s14; do
  if [ $v = default ]; then unset GM_LIB_PATH; else export GM_LIB_PATH=$PWD/oriented_object_detection_b200/lib/variants/$v.so; fi
  python /tmp/sel_stats.py >> gpurun_out/r3a_c3.jsonl 2>> gpurun_out/r3a.err
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-iou --no-extras 2>> gpurun_out/r3a.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'lib': '$v', 'ms_per_step': d['ms_per_step'], 'select_grad': d['roofline']['stages_ms']['select_grad'], 'select_dist': d['roofline']['stages_ms']['select_dist'], 'build_ms': d['roofline']['dtedge_build_ms']}))" >> gpurun_out/r3a_c5.jsonl
done
unset GM_LIB_PATH
cat gpurun_out/r3a_c3.jsonl gpurun_out/r3a_c5.jsonl; tail -3 gpurun_out/r3a.err
