#!/bin/bash
# Round 2, call P: fused small-tile kernels (parity against the large-tile kernels and the oracle), the 128/30 leg.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_pixel.py tests/test_gpu_zy_otsu.py tests/test_gpu_zz_tma.py tests/test_gpu_train.py tests/test_gpu_geom.py -q 2>&1 | tail -15 > gpurun_out/r2p_pytest.log
cat gpurun_out/r2p_pytest.log
cat > /tmp/leg128.py <<'PY'
import json, os, sys, torch
sys.path.insert(0, os.getcwd())
from oriented_object_detection_b200 import ops, synth
dev = torch.device("cuda:0")
plan = ops.make_plan(8192, 8192, 128, 30, device=dev)
m = synth.synthetic_map(8192, 8192, 1000, dev)
out = torch.empty(4 * plan.total_px, dtype=torch.uint8, device=dev)
acc = {}
for _ in range(6):
    _, ms = ops.dtedge_build_timed(m, plan, out=out)
    for k, v in ms.items(): acc[k] = min(acc.get(k, 1e9), v)
print(json.dumps({"fused": os.environ.get("GM_SMALL_FUSED", "1"), "tiles": plan.n, "stages_ms": acc, "build_ms": sum(acc.values()), "checksum": int(out[::4097].to(torch.int64).sum().item())}))
PY
for f in 1 0; do GM_SMALL_FUSED=$f python /tmp/leg128.py >> gpurun_out/r2p_128.jsonl 2>> gpurun_out/r2p.err; done
cat gpurun_out/r2p_128.jsonl; tail -3 gpurun_out/r2p.err
