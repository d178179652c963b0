#!/bin/bash
# Round 2, final single-GPU visit: all GPU tests, ncu launch list + full captures (CUDA-core, DMMA, split form), bench lines
# (default = 64^3 strong-scaling point, --weak = configs[1]), reference arm, configuration sweep.
O=gpurun_out/r2_final; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
B="--weak --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-self-check"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py $B > $O/ncu_launch.log 2>&1; echo "ncu list rc=$?"
cap() { tag=$1; pat=$2; shift 2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$pat" -s 8 -c 3 -o /tmp/$tag -f python bench.py $B "$@" > $O/ncu_$tag.log 2>&1; echo "$tag rc=$?"
  ncu -i /tmp/$tag.ncu-rep --page raw --csv > $O/${tag}_raw.csv 2>/dev/null; ncu -i /tmp/$tag.ncu-rep --page source --csv > $O/${tag}_source.csv 2>/dev/null; }
cap core "k_volume|k_gradient|k_riemann"
H3D_USE_MMA=1 cap dmma "k_volume|k_gradient"
cap split "k_volume|k_riemann" --flow Euler --inviscid split-form --averaging pirozzoli --nodes gauss-lobatto
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?"
timeout 600 python bench.py --weak > $O/bench_configs1.json 2> $O/bench_configs1.err; echo "bench configs1 rc=$?"
timeout 400 python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
run() { tag=$1; shift; timeout 600 python bench.py --weak --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-self-check "$@" > $O/sweep_$tag.json 2> $O/sweep_$tag.err; python - $tag <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2_final/sweep_%s.json'%t).read().strip().splitlines()[-1]); r=d['roofline']
    print("| %s | %s | %.1f M | %.2f | %.2f | %.2f / %.2f / %.2f | %.2f | %s |"%(t,d['config']['workload'].split(',',1)[1].strip().replace(', RK3, fixed dt',''),d['config']['ndof']/1e6,d['value']/1e9,d['ms_per_step'],r['per_kernel_ms']['gradient'],r['per_kernel_ms']['riemann'],r['per_kernel_ms']['volume'],r['stage']['frac'],(d['clocks'] or {}).get('sm_mhz')))
except Exception as ex: print(t,"FAILED",ex, open('gpurun_out/r2_final/sweep_%s.err'%t).read()[-600:])
PY
}
run ns_p7_ne32 > $O/table.md
H3D_USE_MMA=1 run ns_p7_ne32_dmma >> $O/table.md
run c1_ns_p3_ne32 --order 3 >> $O/table.md
run ns_p3_ne64 --ne 64 --order 3 >> $O/table.md
run ns_p5_ne40 --ne 40 --order 5 >> $O/table.md
run euler_std_p7 --flow Euler >> $O/table.md
run c3_split_p3 --flow Euler --inviscid split-form --averaging pirozzoli --nodes gauss-lobatto --ne 64 --order 3 >> $O/table.md
run c3_split_p5 --flow Euler --inviscid split-form --averaging pirozzoli --nodes gauss-lobatto --ne 40 --order 5 >> $O/table.md
run c3_split_p7 --flow Euler --inviscid split-form --averaging pirozzoli --nodes gauss-lobatto >> $O/table.md
run c3_split_p9 --flow Euler --inviscid split-form --averaging pirozzoli --nodes gauss-lobatto --ne 26 --order 9 >> $O/table.md
run ns_split_p7 --inviscid split-form --averaging pirozzoli --nodes gauss-lobatto >> $O/table.md
run c5_les_p3_ne64 --ne 64 --order 3 --les smagorinsky >> $O/table.md
run ns_br2_p7 --viscous BR2 >> $O/table.md
run ns_ip_p7 --viscous IP >> $O/table.md
run ns_energy_p7 --gradient-variables Energy >> $O/table.md
cat $O/table.md
ls -la $O | head -40
