#!/bin/bash
# Round 2: c5 bench at N ranks (seam-band exchange over NCCL), N = $1
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r2d_n$N.json 2> gpurun_out/r2d_n$N.err
echo "rc=$?"; tail -c 2500 gpurun_out/r2d_n$N.err; cat gpurun_out/r2d_n$N.json | cut -c 1-6000
