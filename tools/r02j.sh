#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference_suite.py -q 2>&1 | tail -5
timeout 900 compute-sanitizer --tool synccheck --log-file gpurun_out/r02_sanitizer_synccheck_strip.log python tools/sanitize_target.py all 0 > gpurun_out/r02_sanitizer_synccheck_strip_stdout.log 2>&1
echo "== synccheck strip exit $?"; tail -3 gpurun_out/r02_sanitizer_synccheck_strip_stdout.log; tail -3 gpurun_out/r02_sanitizer_synccheck_strip.log
timeout 900 compute-sanitizer --tool synccheck --log-file gpurun_out/r02_sanitizer_synccheck_vjp8.log python tools/sanitize_target.py vjp 1 > gpurun_out/r02_sanitizer_synccheck_vjp8_stdout.log 2>&1
echo "== synccheck vjp8 exit $?"; tail -3 gpurun_out/r02_sanitizer_synccheck_vjp8_stdout.log; grep -c "Barrier error" gpurun_out/r02_sanitizer_synccheck_vjp8.log; grep "by thread" gpurun_out/r02_sanitizer_synccheck_vjp8.log | sort | uniq -c | head
