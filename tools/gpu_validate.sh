#!/bin/bash
# One validation pass on a B200 box: GPU tests, bench line, ncu launch list, full captures of the step kernel (a real
# 6-stage launch: the third k_step of a forward) and of the callers' kernels.
# Usage (under gpurun): bash tools/gpu_validate.sh <tag>
tag=${1:-r01x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
cat gpurun_out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err; echo "reference arm exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --skip-cpu --train-batch 0 > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 2 -c 1 -f -o gpurun_out/${tag}_k_step \
  python bench.py --steps 1 --warmup 3 --skip-cpu --train-batch 0 > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_convs2|k_resconv|k_stem|k_groupnorm" -c 7 -f -o gpurun_out/${tag}_callers \
  python bench.py --steps 1 --warmup 3 --skip-cpu --train-batch 0 > gpurun_out/${tag}_ncu_callers.log 2>&1; echo "ncu callers exit $?"
ls -la gpurun_out
