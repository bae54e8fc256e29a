#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step8 -s 3 -c 1 -f -o gpurun_out/r02g_k_step8 \
  python bench.py --steps 1 --warmup 3 --skip-cpu --train-batch 0 > gpurun_out/r02g_ncu_step8.log 2>&1; echo "ncu step8 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_vjp8|k_wgrad" -s 6 -c 2 -f -o gpurun_out/r02g_k_vjp8 \
  python tools/train_profile.py 4736 > gpurun_out/r02g_ncu_vjp8.log 2>&1; echo "ncu vjp8 exit $?"
ls -la gpurun_out | tail -5
