#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/sweep_$tag.json 2> gpurun_out/sweep_$tag.err; python - $tag <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/sweep_%s.json'%t).read().strip().splitlines()[-1]); r=d['roofline']
    print("%-22s %7.3f GDOF/s %7.2f ms/step  grad %.3f riem %.3f vol %.3f  stage-frac %.3f  ndof %d"%(t,d['value']/1e9,d['ms_per_step'],r['per_kernel_ms']['gradient'],r['per_kernel_ms']['riemann'],r['per_kernel_ms']['volume'],r['stage']['frac'],d['config']['ndof']))
except Exception as ex: print(t,"FAILED",ex, open('gpurun_out/sweep_%s.err'%t).read()[-800:])
PY
}
run wale_p7_ne32 --order 7 --ne 32 --les wale
run vreman_p7_ne32 --order 7 --ne 32 --les vreman
run vreman_p3_ne64 --order 3 --ne 64 --les vreman
