#!/bin/bash
# Round 2: DMMA passes with several tiles in flight; bench.py with the strong-scaling default (64^3 on one GPU) and the
# single-GPU self-check values for tests/golden/scale_check.json
mkdir -p gpurun_out/r2_f
run() { tag=$1; shift; t0=$SECONDS; timeout 1500 python bench.py "$@" > gpurun_out/r2_f/$tag.json 2> gpurun_out/r2_f/$tag.err; echo "$tag rc=$? wall=$((SECONDS-t0))s"; python - $tag <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2_f/%s.json'%t).read().strip().splitlines()[-1]); r=d['roofline']
    print("%-22s %7.3f GDOF/s %7.2f ms/step  grad %.3f riem %.3f vol %.3f  stage-frac %.3f"%(t,d['value']/1e9,d['ms_per_step'],r['per_kernel_ms']['gradient'],r['per_kernel_ms']['riemann'],r['per_kernel_ms']['volume'],r['stage']['frac']))
    print("   self_check", d.get('self_check'), " e2e", d.get('e2e'), " cpu", d.get('cpu_baseline'))
except Exception as ex: print(t,"FAILED",ex, open('gpurun_out/r2_f/%s.err'%t).read()[-1500:])
PY
}
run weak32_core --weak --steps 30 --no-e2e --no-cpu-baseline
H3D_USE_MMA=1 run weak32_mma --weak --steps 30 --no-e2e --no-cpu-baseline --no-self-check
timeout 600 python scripts/mma_parity.py > gpurun_out/r2_f/mma_parity.txt 2>&1; tail -4 gpurun_out/r2_f/mma_parity.txt
run default_strong64 --steps 20
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_f/reference.json 2> gpurun_out/r2_f/reference.err; cat gpurun_out/r2_f/reference.json
