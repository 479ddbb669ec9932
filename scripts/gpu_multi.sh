#!/bin/bash
# N-GPU bench (torchrun, one rank per GPU over NCCL).  usage: gpu_multi.sh <N>
N=${1:-2}
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-iou > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo rc=$?; cat gpurun_out/bench_n$N.json; tail -15 gpurun_out/bench_n$N.err
