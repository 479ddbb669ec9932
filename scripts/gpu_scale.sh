#!/bin/bash
# c5 bench at N = 1, 2, 4, 8 on one 8-GPU box, back to back (the driver's scaling run).
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_scale_n$N.json 2> gpurun_out/${TAG}_scale_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540 + N)) bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/${TAG}_scale_n$N.json 2> gpurun_out/${TAG}_scale_n$N.err
  fi
  echo "N=$N rc=$?"; tail -c 600 gpurun_out/${TAG}_scale_n$N.err | tail -n 4
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_scale_n$N.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("n_gpus", "value", "ms_per_step")}, d["e2e"]["value"], d["e2e"]["h2d_probe_gbs"], d["config"]["merged_checksum"], d["config"]["seam_rows_exchanged"],
          d["roofline"]["dtedge_build_ms"], d["roofline"]["merge_path_wall_ms"], d["iou"]["gpairs_per_s"] if d.get("iou") else None)
except Exception as e:
    print("no line:", e)
PY
done
