#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
for o in "5 40" "7 32"; do set -- $o
python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --order $1 --ne $2 --flow Euler --inviscid split-form --averaging pirozzoli --nodes gauss-lobatto 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('P$1 %.3f G vol %.3f ms clocks %s'%(d['value']/1e9,d['roofline']['per_kernel_ms']['volume'],d['clocks']['sm_mhz']))"
done; done
