#!/bin/bash
# Round 2, call G: TMA probe (aligned starts), TMA gradient kernel parity + A/B timing + ncu.
set -u
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -DBOXW=160 -o /tmp/tma_probe scripts/probes/tma_probe.cu && timeout 60 /tmp/tma_probe 2>&1 | tee gpurun_out/tma_probe.log
timeout 900 python -m pytest tests/test_gpu_zz_tma.py tests/test_gpu_pixel.py -q -x 2>&1 | tail -8 > gpurun_out/r2g_pytest.log
cat gpurun_out/r2g_pytest.log
cat > /tmp/grad_leg.py <<'PY'
import json, os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from oriented_object_detection_b200 import ops, synth
dev = torch.device("cuda:0")
plan = ops.make_plan(8192, 8192, 416, 100, device=dev)
m = synth.synthetic_map(8192, 8192, 1000, dev)
out = torch.empty(4 * plan.total_px, dtype=torch.uint8, device=dev)
acc = {}
for _ in range(6):
    _, ms = ops.dtedge_build_timed(m, plan, out=out)
    for k, v in ms.items(): acc[k] = min(acc.get(k, 1e9), v)
print(json.dumps({"tma": os.environ.get("GM_GRAD_TMA", "1"), "stages_ms": acc, "build_ms": sum(acc.values()), "checksum": int(out[::4097].to(torch.int64).sum().item())}))
PY
for v in 1 0; do GM_GRAD_TMA=$v python /tmp/grad_leg.py >> gpurun_out/r2g_grad.jsonl 2>> gpurun_out/r2g.err; done
cat gpurun_out/r2g_grad.jsonl; tail -3 gpurun_out/r2g.err
python /tmp/grad_leg.py > gpurun_out/r2g_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_grad_fast' -s 2 -c 1 -o gpurun_out/r2g_prof_grad python /tmp/grad_leg.py > gpurun_out/r2g_ncu2.log 2>&1
tail -n 2 gpurun_out/r2g_ncu2.log | cut -c 1-300
