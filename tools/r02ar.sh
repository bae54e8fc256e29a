#!/bin/bash
python -c "import sys; sys.path.insert(0,'.'); import __graft_entry__ as g; print('stale', g._stale())"
timeout 900 python -m pytest tests/test_gpu_each.py tests/test_gpu_fused.py tests/test_gpu_vjp.py tests/test_gpu_adjoint.py tests/test_gpu_rejects.py tests/test_gpu_unrolled.py -x -q 2>&1 | tail -4
BATCHES=1,2,4,128 timeout 200 python tools/vjp_latency.py 2>&1 | tail -4
