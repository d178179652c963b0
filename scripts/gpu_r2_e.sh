#!/bin/bash
# Round 2: unpadded work fields + warp-wide TMA issue in the gen1 kernels: parity, speed (CUDA-core and DMMA), config sweep
mkdir -p gpurun_out/r2_e
timeout 1200 python -m pytest tests/test_gpu_parity_large.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_e/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_e/pytest.log
run() { tag=$1; shift; timeout 900 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/r2_e/$tag.json 2> gpurun_out/r2_e/$tag.err; python - $tag <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2_e/%s.json'%t).read().strip().splitlines()[-1]); r=d['roofline']
    print("%-22s %7.3f GDOF/s %7.2f ms/step  grad %.3f riem %.3f vol %.3f  stage-frac %.3f"%(t,d['value']/1e9,d['ms_per_step'],r['per_kernel_ms']['gradient'],r['per_kernel_ms']['riemann'],r['per_kernel_ms']['volume'],r['stage']['frac']))
except Exception as ex: print(t,"FAILED",ex, open('gpurun_out/r2_e/%s.err'%t).read()[-800:])
PY
}
run ns_p7_core
H3D_USE_MMA=1 run ns_p7_mma
run ns_p3_ne64 --ne 64 --order 3
run ns_p5_ne40 --ne 40 --order 5
run split_p7 --flow Euler --inviscid split-form --averaging pirozzoli --nodes gauss-lobatto
run split_p3 --flow Euler --inviscid split-form --averaging pirozzoli --nodes gauss-lobatto --ne 64 --order 3
run split_p5 --flow Euler --inviscid split-form --averaging pirozzoli --nodes gauss-lobatto --ne 40 --order 5
run split_p9 --flow Euler --inviscid split-form --averaging pirozzoli --nodes gauss-lobatto --ne 26 --order 9
run ns_split_p7 --inviscid split-form --averaging pirozzoli --nodes gauss-lobatto
run euler_std_p7 --flow Euler
