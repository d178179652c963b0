#!/bin/bash
# ncu --set full capture of the three hot kernels (one launch each, taken in steady state)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_volume|k_gradient|k_riemann" -s 9 -c 3 -o gpurun_out/hot_full -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline "$@" > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
