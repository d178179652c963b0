#!/bin/bash
# quick GPU visit: parity tests + one short bench line (kernel numbers only)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
scripts/bench_env.sh "${@:-H3D_NOOP=1}"
