#!/bin/bash
# training step at batch 4736 with the step-wise adjoint route (the kernels of the conditional-graph loop are not visible to ncu)
mkdir -p gpurun_out
export NODE_B200_ADJOINT_SOLVE=0
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02bg_train_step_launches_b4736.csv python tools/train_profile.py 4736 > gpurun_out/r02bg_train.log 2>&1; echo "launch list exit $?"
tail -2 gpurun_out/r02bg_train.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_vjp8|k_wgrad8" -s 6 -c 2 -f -o gpurun_out/r02bg_vjp8_wgrad8 python tools/train_profile.py 4736 > gpurun_out/r02bg_ncu_full.log 2>&1; echo "full exit $?"
