#!/bin/bash
# Round 2, call S: full GPU suite on the final library, full bench line, reference arm, c3 / 128-30 stage legs, smoke.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2s_pytest.log
cat gpurun_out/r2s_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python scripts/probes/grad_leg.py > gpurun_out/r2s_c3.json 2>> gpurun_out/r2s.err; cat gpurun_out/r2s_c3.json
timeout 900 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2s_bench.err
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2s_bench_ref.json 2> gpurun_out/r2s_bench_ref.err; echo "ref rc=$?"
python -c "
import json
d = json.loads(open('gpurun_out/r2s_bench.json').read().strip().splitlines()[-1])
print(json.dumps({'value': d['value'], 'ms_per_step': d['ms_per_step'], 'e2e': d['e2e']['value'], 'launches': d['gpu_launches'], 'roofline': {k: d['roofline'][k] for k in ('frac', 'dtedge_build_ms', 'dtedge_build_frac_of_hbm', 'stages_ms', 'merge_path_wall_ms', 'tile_gather3_frac_of_hbm')}, 'iou': d['iou']['gpairs_per_s'], 'frac_nominal': d['iou']['frac_of_nominal_fp32'], 'dtedge_128': d.get('dtedge_128'), 'fusion_ms': d['fusion']['fusion_ms'], 'cpu': d['cpu_baseline']['value']}))
r = json.loads(open('gpurun_out/r2s_bench_ref.json').read().strip().splitlines()[-1]); print(r['value'], r['cpu_baseline']['kind'], r['cpu_baseline']['cores'])"
