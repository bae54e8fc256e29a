#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py -x -q 2>&1 | tail -5
NS=128 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r02an_wide_vjp_c256_b128.csv python tools/wide_vjp_debug.py > gpurun_out/r02an_ncu.log 2>&1
tail -2 gpurun_out/r02an_ncu.log
