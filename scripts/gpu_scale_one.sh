#!/bin/bash
# c5 bench at ONE N on a box with N GPUs (gpurun --gpus N): usage gpu_scale_one.sh N TAG
set -u
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-r2b}
if [ $N -eq 1 ]; then
  timeout 900 python bench.py --gpus 1 --steps 8 --warmup 3 --no-extras > gpurun_out/${TAG}_scale_n$N.json 2> gpurun_out/${TAG}_scale_n$N.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540 + N)) bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/${TAG}_scale_n$N.json 2> gpurun_out/${TAG}_scale_n$N.err
fi
echo "N=$N rc=$?"; tail -c 600 gpurun_out/${TAG}_scale_n$N.err | tail -n 4
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_scale_n$N.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("n_gpus", "value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e"]["h2d_probe_gbs"], d["config"]["merged_checksum"], "seam rows", d["config"]["seam_rows_exchanged"],
          "fallbacks", d["config"].get("seam_chain_fallbacks_rank0"), "build", d["roofline"]["dtedge_build_ms"], "merge", d["roofline"]["merge_path_wall_ms"], "iou", d["iou"]["gpairs_per_s"] if d.get("iou") else None, d["clocks"])
except Exception as e:
    print("no line:", e)
PY
