#!/bin/bash
# Round 2, first visit: toolchain probe, DMMA micro-benchmark, multi-tile / full-size per-DOF parity tests.
mkdir -p gpurun_out/r2_a
{ echo "gfortran: $(which gfortran || echo none)"; echo "mpirun: $(which mpirun mpiexec mpif90 || echo none)"; echo "nproc: $(nproc)"; lscpu | grep -E "Model name|Socket|Core|Thread"; free -g | head -2; nvidia-smi -L; } > gpurun_out/r2_a/probe.txt 2>&1
cat gpurun_out/r2_a/probe.txt
./scripts/micro/dmma_rate > gpurun_out/r2_a/dmma_rate.txt 2>&1; cat gpurun_out/r2_a/dmma_rate.txt
timeout 1500 python -m pytest tests/test_gpu_parity_large.py -m gpu -x -q -s > gpurun_out/r2_a/pytest_large.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_a/pytest_large.log
