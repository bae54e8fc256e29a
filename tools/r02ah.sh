#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unrolled.py tests/test_gpu_reference_suite.py tests/test_gpu_adjoint.py -x -q 2>&1 | tail -15
