#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/vjp_latency.py 2>&1 | tail -4
REPS=5 BATCHES=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r02ag_vjp_b1_launches.csv python tools/vjp_latency.py > gpurun_out/r02ag_ncu.log 2>&1
tail -2 gpurun_out/r02ag_ncu.log
