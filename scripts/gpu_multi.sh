#!/bin/bash
# multi-GPU visit: 2-rank parity test + weak-scaling bench line at N ranks (N = first argument)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q > gpurun_out/pytest_multirank.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_multirank.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_n$N.json
