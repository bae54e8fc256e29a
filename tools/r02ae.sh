#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_adjoint.py tests/test_gpu_vjp.py tests/test_gpu_reference_suite.py -x -q 2>&1 | tail -15
timeout 300 python tools/pgd_latency.py 2>&1 | tail -12
