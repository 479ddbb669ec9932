#!/bin/bash
# N-GPU bench (torchrun, one rank per GPU over NCCL).  usage: gpu_multi.sh <N> [graph modes...]
N=${1:-2}; shift
MODES=${@:-auto}
mkdir -p gpurun_out
for G in $MODES; do
  timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-iou --graph $G > gpurun_out/bench_n${N}_$G.json 2> gpurun_out/bench_n${N}_$G.err
  echo "graph=$G rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n${N}_$G.json").read().strip().splitlines()[-1])
    print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"]["detection_path"], d["roofline"]["merge_path_wall_ms"], d["config"]["merged"])
except Exception as e:
    print("no json:", e)
PY
  tail -4 gpurun_out/bench_n${N}_$G.err | cut -c1-300
done
