#!/bin/bash
# Round 2, call B: seam-band exchange kernels (emulated ranks) + the full GPU suite.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_seam.py -q -x 2>&1 | tail -15 > gpurun_out/r2b_seam.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2b_pytest.log
cat gpurun_out/r2b_seam.log; tail -6 gpurun_out/r2b_pytest.log
