#!/bin/bash
# Round 2, closing visit: every single-GPU test on the final library and the ncu capture bench.py quotes (tagged with the source hash)
O=gpurun_out/r2_q; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
B="--weak --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-self-check"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_volume|k_gradient|k_riemann" -s 9 -c 3 -o /tmp/core -f python bench.py $B > $O/ncu_core.log 2>&1; echo "core rc=$?"
ncu -i /tmp/core.ncu-rep --page raw --csv > $O/core_raw.csv 2>/dev/null; ncu -i /tmp/core.ncu-rep --page source --csv > $O/core_source.csv 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
