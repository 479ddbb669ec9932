#!/bin/bash
# selection variants: parity tests of the pixel path + stage timing with and without the sampled front end
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pixel.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_pixel.log
cat gpurun_out/pytest_pixel.log
timeout 200 python scripts/sweep_select.py 2>&1 | tail -12
