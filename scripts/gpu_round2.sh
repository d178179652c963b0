#!/bin/bash
# Second GPU-box visit of the round: all parity tests, the strong-scaling 1-GPU point (64^3 P7 on one GPU), bench lines for
# the BR2 / IP / entropy-variable variants, the full bench line with the reference arm, ncu launch list and full capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
run() { tag=$1; shift; timeout 900 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/sweep_$tag.json 2> gpurun_out/sweep_$tag.err; python - $tag <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/sweep_%s.json'%t).read().strip().splitlines()[-1]); r=d['roofline']
    print("%-22s %7.3f GDOF/s %7.2f ms/step  grad %.3f riem %.3f vol %.3f  stage-frac %.3f  ndof %d"%(t,d['value']/1e9,d['ms_per_step'],r['per_kernel_ms']['gradient'],r['per_kernel_ms']['riemann'],r['per_kernel_ms']['volume'],r['stage']['frac'],d['config']['ndof']))
except Exception as ex: print(t,"FAILED",ex, open('gpurun_out/sweep_%s.err'%t).read()[-800:])
PY
}
run ns_br2_p7 --viscous BR2
run ns_ip_p7 --viscous IP
run ns_entropy_split_p7 --inviscid split-form --averaging chandrasekar --riemann "matrix dissipation" --nodes gauss-lobatto --gradient-variables Entropy
run ns_energy_p7 --gradient-variables Energy
avail=$(awk '/MemAvailable/{print int($2/1048576)}' /proc/meminfo); echo "host MemAvailable ${avail} GiB, cores $(nproc)"
if [ "$avail" -ge 150 ]; then run strong_1gpu_ne64_p7 --ne 64 --steps 10; else echo "strong_1gpu_ne64_p7 skipped: not enough host memory"; fi
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_full.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cat gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_volume|k_gradient|k_riemann" -s 9 -c 3 -o gpurun_out/hot_full -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | head -40
