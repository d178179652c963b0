#!/bin/bash
# Round 2, two GPUs: MPI-face geometry taken from the left-side owner: the "local geometry" cases must now reproduce the single-domain
# oracle to the last bits; strong-scaling point N=2 with the self-check at 1e-11
mkdir -p gpurun_out/r2_p
timeout 500 python -m pytest tests/test_gpu_multirank.py tests/test_cylinder_tutorial.py -m gpu -q -s -k "(w2 and local) or w2-ne4-N3-NS-metis-inherit or 2]" > gpurun_out/r2_p/pytest.log 2>&1; echo "pytest rc=$?"; grep "multirank\|configs\|passed\|failed\|Error" gpurun_out/r2_p/pytest.log | tail -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/r2_p/bench_n2.json 2> gpurun_out/r2_p/bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_p/bench_n2.json').read().strip().splitlines() if l.startswith('{')][-1])
    print("N=2 %.3f GDOF/s %.2f ms/step self_check %s"%(d['value']/1e9,d['ms_per_step'],d['self_check']))
    for r in d['timeline']['ms_per_rank']: print("   ", r)
except Exception as ex: print("FAILED", ex, open('gpurun_out/r2_p/bench_n2.err').read()[-1500:])
PY
