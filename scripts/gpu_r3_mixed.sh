#!/bin/bash
# First GPU visit of the p-nonconforming path: parity on the device, throughput, launch list, one full ncu capture.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_r3_mixed.sh'
set -x
OUT=gpurun_out/r3_mixed; mkdir -p $OUT
python -m pytest tests/test_zz_gpu_mixed.py -q -m gpu -s > $OUT/pytest_mixed.log 2>&1; tail -3 $OUT/pytest_mixed.log
# with gpurun --gpus 2|4|8: the partitioned cases (NCCL exchange of the traces at the face order)
python -m pytest tests/test_zz_gpu_mixed_multirank.py -q -m gpu -s > $OUT/pytest_mixed_multirank.log 2>&1; tail -3 $OUT/pytest_mixed_multirank.log
compute-sanitizer --tool memcheck python -m pytest tests/test_zz_gpu_mixed.py -q -m gpu -k "golden_mixed or les_models" > $OUT/sanitizer_memcheck.log 2>&1; tail -3 $OUT/sanitizer_memcheck.log
python scripts/bench_mixed.py --ne 24 --lo 3 --hi 7 --steps 20 --warmup 3 > $OUT/bench_mixed.json 2> $OUT/bench_mixed.err; cat $OUT/bench_mixed.json
python scripts/bench_mixed.py --ne 32 --lo 7 --hi 7 --steps 10 --warmup 3 > $OUT/bench_mixed_uniform_p7.json 2>> $OUT/bench_mixed.err; cat $OUT/bench_mixed_uniform_p7.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv python scripts/bench_mixed.py --ne 16 --lo 3 --hi 7 --steps 2 --warmup 1 > $OUT/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mx -c 13 -s 52 -o $OUT/mixed_full python scripts/bench_mixed.py --ne 16 --lo 3 --hi 7 --steps 2 --warmup 1 > $OUT/ncu_full.log 2>&1
