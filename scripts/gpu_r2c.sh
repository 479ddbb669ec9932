#!/bin/bash
# Round 2, call C: first runs of the c5 bench (N=1): small then default.
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --maps 2 --no-cpu-baseline --no-extras > gpurun_out/r2c_small.json 2> gpurun_out/r2c_small.err
echo "small rc=$?"; tail -c 1500 gpurun_out/r2c_small.err
timeout 1500 python bench.py --steps 8 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
echo "full rc=$?"; tail -c 3000 gpurun_out/r2c_bench.err
cat gpurun_out/r2c_small.json gpurun_out/r2c_bench.json
