#!/bin/bash
# Round 2, two GPUs: multi-rank parity (2-rank cases) and the strong-scaling point N=2 with the stage timeline
mkdir -p gpurun_out/r2_h
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -s -k "w2" > gpurun_out/r2_h/pytest_multirank.log 2>&1; echo "pytest rc=$?"; grep "multirank\|passed\|failed\|Error" gpurun_out/r2_h/pytest_multirank.log | tail -24
for sms in 8 0; do
H3D_COMM_SMS=$sms timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/r2_h/bench_n2_sms$sms.json 2> gpurun_out/r2_h/bench_n2_sms$sms.err; echo "bench n2 comm_sms=$sms rc=$?"
python - $sms <<'PY'
import json,sys
try:
    d=json.loads([l for l in open('gpurun_out/r2_h/bench_n2_sms%s.json'%sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    print("comm_sms=%s N=2 %.3f GDOF/s %.2f ms/step self_check %s"%(sys.argv[1],d['value']/1e9,d['ms_per_step'],d['self_check']))
    for r in d['timeline']['ms_per_rank']: print("   ", r)
except Exception as ex: print("FAILED", ex, open('gpurun_out/r2_h/bench_n2_sms%s.err'%sys.argv[1]).read()[-1500:])
PY
done
