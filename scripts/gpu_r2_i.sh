#!/bin/bash
# Round 2: the whole single-GPU test suite on the current library (incl. the K4 / K5 regressions and configs[4] on the device)
mkdir -p gpurun_out/r2_i
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r2_i/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_i/pytest_gpu.log
