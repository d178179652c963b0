#!/bin/bash
# Round 2, after the removal of the dead two-nodes-per-thread configuration (generated code unchanged): parity again, the ncu
# capture bench.py quotes (tagged with the hash of these sources), launch list
O=gpurun_out/r2_m; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity_large.py tests/test_gpu_parity.py tests/test_golden.py tests/test_cpp_driver.py -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
B="--weak --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-self-check"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py $B > $O/ncu_launch.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_volume|k_gradient|k_riemann" -s 8 -c 3 -o /tmp/core -f python bench.py $B > $O/ncu_core.log 2>&1; echo "core rc=$?"
ncu -i /tmp/core.ncu-rep --page raw --csv > $O/core_raw.csv 2>/dev/null; ncu -i /tmp/core.ncu-rep --page source --csv > $O/core_source.csv 2>/dev/null
ls -la $O
