#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_case.py 7 > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done|Error|error" gpurun_out/sanitize_$tool.log | head -8
done
