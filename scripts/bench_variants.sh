#!/bin/bash
# usage: scripts/bench_variants.sh lib1.so lib2.so ...   (run on the GPU box; prints one short line per variant)
for lib in "$@"; do
  H3D_GPU_LIB=$PWD/horses3d_b200/csrc/$lib python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_$lib.json 2> gpurun_out/bench_$lib.err
  python - "$lib" <<'PY'
import json,sys
lib=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/bench_%s.json'%lib).read().strip().splitlines()[-1])
    r=d['roofline']
    print("%-28s %.3f GDOF/s  %.2f ms/step  grad %.3f  riem %.3f  vol %.3f ms  stage-frac %.3f"%(lib,d['value']/1e9,d['ms_per_step'],r['per_kernel_ms']['gradient'],r['per_kernel_ms']['riemann'],r['per_kernel_ms']['volume'],r['stage']['frac']))
except Exception as ex:
    print(lib,"FAILED",ex); print(open('gpurun_out/bench_%s.err'%lib).read()[-800:])
PY
done
