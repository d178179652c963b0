#!/bin/bash
# ncu --set full capture of the hot kernels of one bench.py configuration, exported to CSV on the box.
# usage: scripts/gpu_ncu.sh <tag> <kernel regex> [bench.py args...]     (environment H3D_* is passed through)
tag=$1; pat=$2; shift 2
mkdir -p gpurun_out/ncu
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$pat" -s 8 -c 3 -o /tmp/$tag -f python bench.py --weak --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-self-check "$@" > gpurun_out/ncu/$tag.log 2>&1; echo "$tag rc=$?"
ncu -i /tmp/$tag.ncu-rep --page raw --csv > gpurun_out/ncu/${tag}_raw.csv 2>/dev/null
ncu -i /tmp/$tag.ncu-rep --page source --csv > gpurun_out/ncu/${tag}_source.csv 2>/dev/null
ls -la gpurun_out/ncu | tail -4
