#!/bin/bash
# Round-2 state check: GPU tests, dense-step timeline (debug build), training-step launch list at the bench batch.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; tail -3 gpurun_out/r02d_pytest.log
NODE_B200_LIB=gpurun_variants/libnode_b200_dbg.so timeout 300 python tools/step8_timeline.py 4736 > gpurun_out/r02d_step8_timeline.txt 2>&1; cat gpurun_out/r02d_step8_timeline.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02d_train_launches.csv python tools/train_profile.py 4736 > gpurun_out/r02d_train_profile.log 2>&1; tail -3 gpurun_out/r02d_train_profile.log
