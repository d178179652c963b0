#!/bin/bash
# Round 2: second-generation n=8 kernels (2 CTAs/SM): parity, then speed of the four variants.
mkdir -p gpurun_out/r2_c
timeout 900 python -m pytest tests/test_gpu_parity_large.py -m gpu -x -q -k "8-7 or config1 or boundary" > gpurun_out/r2_c/pytest_large.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_c/pytest_large.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_c/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_c/pytest_parity.log
timeout 600 python scripts/mma_parity.py > gpurun_out/r2_c/mma_parity.txt 2>&1; tail -5 gpurun_out/r2_c/mma_parity.txt
for cfg in "0 0" "1 0" "1 1" "0 1"; do set -- $cfg
H3D_GEN2=$1 H3D_USE_MMA=$2 timeout 600 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_c/bench_gen$1_mma$2.json 2> gpurun_out/r2_c/bench_gen$1_mma$2.err; echo "bench gen2=$1 mma=$2 rc=$?"
python - $1 $2 <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2_c/bench_gen%s_mma%s.json'%(sys.argv[1],sys.argv[2])).read().strip().splitlines()[-1]); r=d['roofline']
print("gen2=%s mma=%s %7.3f GDOF/s %7.2f ms/step  grad %.3f riem %.3f vol %.3f  stage-frac %.3f"%(sys.argv[1],sys.argv[2],d['value']/1e9,d['ms_per_step'],r['per_kernel_ms']['gradient'],r['per_kernel_ms']['riemann'],r['per_kernel_ms']['volume'],r['stage']['frac']))
PY
done
