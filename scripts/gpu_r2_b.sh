#!/bin/bash
# Round 2, DMMA experiment: error against the oracle and speed of both contraction paths.
mkdir -p gpurun_out/r2_b
timeout 900 python scripts/mma_parity.py 16 > gpurun_out/r2_b/mma_parity.txt 2>&1; cat gpurun_out/r2_b/mma_parity.txt | tail -8
for mma in 0 1; do
H3D_USE_MMA=$mma timeout 600 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_b/bench_mma$mma.json 2> gpurun_out/r2_b/bench_mma$mma.err; echo "bench mma=$mma rc=$?"
python - $mma <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2_b/bench_mma%s.json'%sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print("mma=%s %7.3f GDOF/s %7.2f ms/step  grad %.3f riem %.3f vol %.3f  stage-frac %.3f"%(sys.argv[1],d['value']/1e9,d['ms_per_step'],r['per_kernel_ms']['gradient'],r['per_kernel_ms']['riemann'],r['per_kernel_ms']['volume'],r['stage']['frac']))
PY
done
