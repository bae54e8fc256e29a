#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_wide.py -x -q -s 2>&1 | tail -12
timeout 600 python tools/wide_train_bench.py 2>&1 | tail -8
