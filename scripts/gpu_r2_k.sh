#!/bin/bash
# Round 2, eight GPUs: 8-rank parity cases, configs[4] on 8 ranks, strong-scaling point N=8 of the 64^3 P=7 mesh with the stage timeline
mkdir -p gpurun_out/r2_k
nvidia-smi -L | wc -l
timeout 420 python -m pytest tests/test_gpu_multirank.py tests/test_cylinder_tutorial.py -m gpu -q -s -k "w8 or partitioned-8 or 8]" > gpurun_out/r2_k/pytest.log 2>&1; echo "pytest rc=$?"; grep "multirank\|configs\|passed\|failed\|Error" gpurun_out/r2_k/pytest.log | tail -12
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e > gpurun_out/r2_k/bench_n8.json 2> gpurun_out/r2_k/bench_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_k/bench_n8.json').read().strip().splitlines() if l.startswith('{')][-1])
    print("N=8 %.3f GDOF/s %.2f ms/step self_check %s"%(d['value']/1e9,d['ms_per_step'],d['self_check']))
    for r in d['timeline']['ms_per_rank']: print("   ", r)
except Exception as ex: print("FAILED", ex, open('gpurun_out/r2_k/bench_n8.err').read()[-1500:])
PY
