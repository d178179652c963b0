#!/bin/bash
# Round 2, two GPUs: remaining 2-rank cases, configs[4] on two ranks, strong-scaling point N=2 with / without the interior split
mkdir -p gpurun_out/r2_j
timeout 600 python -m pytest tests/test_gpu_multirank.py tests/test_cylinder_tutorial.py -m gpu -q -s -k "w2-ne6-N7-NS-metis-local or w2-ne4-N3-IP-metis-local or partitioned" > gpurun_out/r2_j/pytest.log 2>&1; echo "pytest rc=$?"; grep "multirank\|configs\|passed\|failed\|Error" gpurun_out/r2_j/pytest.log | tail -12
for split in 70 0; do
H3D_INTERIOR_SPLIT_PCT=$split timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/r2_j/bench_n2_split$split.json 2> gpurun_out/r2_j/bench_n2_split$split.err; echo "bench n2 split=$split rc=$?"
python - $split <<'PY'
import json,sys
try:
    d=json.loads([l for l in open('gpurun_out/r2_j/bench_n2_split%s.json'%sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    print("split=%s N=2 %.3f GDOF/s %.2f ms/step self_check %s"%(sys.argv[1],d['value']/1e9,d['ms_per_step'],d['self_check']))
    for r in d['timeline']['ms_per_rank']: print("   ", r)
except Exception as ex: print("FAILED", ex, open('gpurun_out/r2_j/bench_n2_split%s.err'%sys.argv[1]).read()[-1500:])
PY
done
