#!/bin/bash
python -c "import sys; sys.path.insert(0,'.'); import __graft_entry__ as g; print('stale', g._stale())"
timeout 300 python -m pytest tests/test_gpu_adjoint.py -x -q 2>&1 | tail -12
