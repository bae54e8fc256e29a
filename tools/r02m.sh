#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02m_launches.csv \
  python bench.py --steps 2 --warmup 3 --skip-cpu --train-batch 0 > gpurun_out/r02m_ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_convs2|k_resconv|k_stem" -c 5 -f -o gpurun_out/r02m_callers \
  python bench.py --steps 1 --warmup 3 --skip-cpu --train-batch 0 > gpurun_out/r02m_ncu_callers.log 2>&1; echo "ncu callers exit $?"
