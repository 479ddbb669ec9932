#!/bin/bash
set -u
mkdir -p gpurun_out
for bw in 144 128 160; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -DBOXW=$bw -o /tmp/tma_probe_$bw scripts/probes/tma_probe.cu && timeout 60 /tmp/tma_probe_$bw
done 2>&1 | tee gpurun_out/tma_probe.log
