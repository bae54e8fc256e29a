#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference_suite.py -x -q 2>&1 | tail -25
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/r02_sanitizer_$tool.log python tools/sanitize_target.py > gpurun_out/r02_sanitizer_${tool}_stdout.log 2>&1
  echo "== $tool exit $?"; tail -3 gpurun_out/r02_sanitizer_${tool}_stdout.log; tail -4 gpurun_out/r02_sanitizer_$tool.log
done
