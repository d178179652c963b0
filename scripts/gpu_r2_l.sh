#!/bin/bash
# Round 2: fused-axis prolongation: parity, then speed (weak 32^3 = configs[1])
mkdir -p gpurun_out/r2_l
timeout 900 python -m pytest tests/test_gpu_parity_large.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_l/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_l/pytest.log
run() { tag=$1; shift; timeout 600 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-self-check "$@" > gpurun_out/r2_l/$tag.json 2> gpurun_out/r2_l/$tag.err; python - $tag <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2_l/%s.json'%t).read().strip().splitlines()[-1]); r=d['roofline']
    print("%-22s %7.3f GDOF/s %7.2f ms/step  grad %.3f riem %.3f vol %.3f  stage-frac %.3f"%(t,d['value']/1e9,d['ms_per_step'],r['per_kernel_ms']['gradient'],r['per_kernel_ms']['riemann'],r['per_kernel_ms']['volume'],r['stage']['frac']))
except Exception as ex: print(t,"FAILED",ex, open('gpurun_out/r2_l/%s.err'%t).read()[-800:])
PY
}
run ns_p7 --weak
run ns_p5 --weak --ne 40 --order 5
run split_p7 --weak --flow Euler --inviscid split-form --averaging pirozzoli --nodes gauss-lobatto
