#!/bin/bash
# Round 2, four GPUs: 4-rank parity cases and configs[4] on 4 ranks
mkdir -p gpurun_out/r2_o
timeout 500 python -m pytest tests/test_gpu_multirank.py tests/test_cylinder_tutorial.py -m gpu -q -s -k "w4 or 4]" > gpurun_out/r2_o/pytest.log 2>&1; echo "pytest rc=$?"; grep "multirank\|configs\|passed\|failed\|Error" gpurun_out/r2_o/pytest.log | tail -8
