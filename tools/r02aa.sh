#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wide.py -x -q 2>&1 | tail -15
timeout 600 python tools/wide_bench.py 2>&1 | tail -8
