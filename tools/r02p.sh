#!/bin/bash
# Full validation of the current tree: GPU tests, default bench line, reference arm.
tag=${1:-r02p}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err; echo "reference arm exit $?"
head -c 1500 gpurun_out/${tag}_bench.json
