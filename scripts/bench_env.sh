#!/bin/bash
# usage: scripts/bench_env.sh "VAR=val ..." "VAR=val ..." : one bench line per environment setting
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_env$i.json 2> gpurun_out/bench_env$i.err
  python - "$envs" $i <<'PY'
import json,sys
tag,i=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open('gpurun_out/bench_env%s.json'%i).read().strip().splitlines()[-1])
    r=d['roofline']
    print("%-28s %.3f GDOF/s  %.2f ms/step  grad %.3f  riem %.3f  vol %.3f ms  stage-frac %.3f"%(tag,d['value']/1e9,d['ms_per_step'],r['per_kernel_ms']['gradient'],r['per_kernel_ms']['riemann'],r['per_kernel_ms']['volume'],r['stage']['frac']))
except Exception as ex:
    print(tag,"FAILED",ex); print(open('gpurun_out/bench_env%s.err'%i).read()[-1500:])
PY
done
