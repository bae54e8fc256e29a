#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vjp.py tests/test_gpu_adjoint.py -x -q 2>&1 | tail -15
timeout 300 python tools/train_profile.py 4736 2>&1 | tail -2
