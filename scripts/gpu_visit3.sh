#!/bin/bash
# parity tests, full bench line (resident e2e with cached monitor passes), compute-sanitizer on the extended case
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print("value %.3f G  e2e %.3f G  resident %.3f G  cpu %.1f M"%(d['value']/1e9,d['e2e']['value']/1e9,d['e2e']['resident']['value']/1e9,d['cpu_baseline']['value']/1e6))
PY
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_case.py 5 > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done|Error|error" gpurun_out/sanitize_$tool.log | head -8
done
