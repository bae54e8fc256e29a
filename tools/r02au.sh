#!/bin/bash
python -c "import sys; sys.path.insert(0,'.'); import __graft_entry__ as g; print('stale', g._stale())"
timeout 600 python -m pytest tests/test_gpu_unrolled.py tests/test_gpu_fused.py tests/test_gpu_reference_suite.py -x -q 2>&1 | tail -12
