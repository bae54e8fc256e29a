#!/bin/bash
python -c "import sys; sys.path.insert(0,'.'); import __graft_entry__ as g; print('stale', g._stale())"
timeout 600 python -m pytest tests/test_gpu_adjoint.py tests/test_gpu_vjp.py tests/test_gpu_reference_suite.py tests/test_gpu_fused.py -x -q 2>&1 | tail -4
cat > /tmp/ab2.py <<'PY'
import os, sys, time
sys.path.insert(0, '.')
import __graft_entry__ as ge
ge.build()
import torch
from node_b200 import models, solver
dev = 'cuda:0'
for B in (1, 128):
    torch.manual_seed(0)
    net = models.ODENet(3, n_filters=64, downsample='residual', tol=1e-3, adjoint=True, dropout=0.5).train().to(dev)
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    x = torch.rand(B, 3, 32, 32, device=dev); y = torch.randint(0, 10, (B,), device=dev)
    def step():
        opt.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(net(x), y).backward()
        opt.step()
    for _ in range(3): step()
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): step()
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / 10 * 1e3)
    print('overlap', os.environ.get('NODE_B200_ADJOINT_OVERLAP', '1'), 'batch', B, 'adjoint train step: %.2f ms (best of 3)' % best, flush=True)
PY
for m in 1 0 1 0; do NODE_B200_ADJOINT_OVERLAP=$m timeout 200 python /tmp/ab2.py 2>&1 | tail -2; done
