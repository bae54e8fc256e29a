#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_adjoint.py -x -q 2>&1 | tail -5
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02af_pgd_b1_launches.csv python tools/pgd_one.py > gpurun_out/r02af_ncu.log 2>&1
tail -3 gpurun_out/r02af_ncu.log
