#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_retrieval.py -x -q 2>&1 | tail -8
python - <<'PY'
import sys, torch, time
sys.path.insert(0, 'neural-ode-features_b200')
import __graft_entry__ as e; e.build()
from node_b200 import retrieval
f = torch.rand(10000, 64, device='cuda')
q, _ = retrieval.normalize_features(f)
for _ in range(3): s = retrieval.retrieval_scores(q, q)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): s = retrieval.retrieval_scores(q, q)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print('scores 10k x 10k x 64: %.3f ms  %.1f GB/s written  %.1f TFLOP/s' % (ms, 0.4 / ms * 1e3, 12.8 / ms))
PY
