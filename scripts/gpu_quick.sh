#!/bin/bash
# quick GPU check: parity tests + bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench.err
tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_quick.json; tail -5 gpurun_out/bench.err
