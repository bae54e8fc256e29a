#!/bin/bash
# One validation pass on a B200 box: GPU tests, bench line, ncu launch list, one full capture of the step kernel.
# Usage (under gpurun): bash tools/gpu_validate.sh <tag>
tag=${1:-r01x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
cat gpurun_out/${tag}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 8 -c 1 -f -o gpurun_out/${tag}_k_step \
  python bench.py --steps 1 --warmup 3 --skip-cpu > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out
