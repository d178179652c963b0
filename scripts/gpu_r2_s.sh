#!/bin/bash
# the bench exactly as the driver launches it on two GPUs (reference arm first), and the single-GPU self-check
O=gpurun_out/r2_s; mkdir -p $O
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > $O/ref_n2.json 2> $O/ref_n2.err; echo "ref rc=$?"; cut -c1-300 $O/ref_n2.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_s/bench_n2.json').read().strip().splitlines() if l.startswith('{')][-1])
    print("N=2 %.3f GDOF/s %.2f ms/step"%(d['value']/1e9,d['ms_per_step'])); print(" self_check", {k:v for k,v in d['self_check'].items() if k!='values'}); print(" e2e", d['e2e']); print(" clocks", d['clocks'], "launches", d['gpu_launches'])
except Exception as ex: print("FAILED", ex, open('gpurun_out/r2_s/bench_n2.err').read()[-1500:])
PY
