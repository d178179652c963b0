#!/bin/bash
# ncu capture of one FIRST-stage launch of each hot kernel (the last stage of a step also stores QDot: 40 B/DOF more)
O=gpurun_out/r2_n; mkdir -p $O
B="--weak --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-self-check"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_volume|k_gradient|k_riemann" -s 9 -c 3 -o /tmp/core -f python bench.py $B > $O/ncu_core.log 2>&1; echo "core rc=$?"
ncu -i /tmp/core.ncu-rep --page raw --csv > $O/core_raw.csv 2>/dev/null; ncu -i /tmp/core.ncu-rep --page source --csv > $O/core_source.csv 2>/dev/null
ls -la $O
