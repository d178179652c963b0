#!/bin/bash
# after the host-geometry change: golden-fixture tests on the device and the self-check values of both bench meshes
O=gpurun_out/r2_r; mkdir -p $O
timeout 600 python -m pytest tests/test_golden.py tests/test_gpu_parity_large.py -m gpu -q -k "golden or config1 or 8-7" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for w in "--weak" ""; do
timeout 600 python bench.py $w --steps 20 --no-e2e --no-cpu-baseline > $O/bench$w.json 2> $O/bench$w.err; echo "bench $w rc=$?"
python - "$w" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2_r/bench%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d['value']/1e9, d['ms_per_step'], d['self_check'], d['roofline']['traffic'], d['clocks'])
PY
done
H3D_DUMP=1 python - <<'PY'
# the values themselves, for tests/golden/scale_check.json
import json,subprocess,sys
PY
