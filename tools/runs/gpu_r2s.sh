#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/sanitize.py 2>&1 | tail -14
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -c "=========" gpurun_out/r02_sanitizer_$tool.log; tail -4 gpurun_out/r02_sanitizer_$tool.log | cut -c1-200
done
