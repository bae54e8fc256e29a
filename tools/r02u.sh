#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vjp.py tests/test_gpu_adjoint.py tests/test_gpu_caller_grad.py -x -q 2>&1 | tail -5
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02u_train_launches.csv python tools/train_profile.py 4736 > gpurun_out/r02u_train_profile.log 2>&1; tail -3 gpurun_out/r02u_train_profile.log
python tools/launch_agg.py gpurun_out/r02u_train_launches.csv 14
