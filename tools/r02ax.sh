#!/bin/bash
python -c "import sys; sys.path.insert(0,'.'); import __graft_entry__ as g; print('stale', g._stale())"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python - <<'PY'
import os, sys, time
sys.path.insert(0, '.')
import __graft_entry__ as ge
ge.build()
import torch
from node_b200 import models, solver
dev = 'cuda:0'
for mode in ('nodes', 'aten', 'adjoint'):
    os.environ['NODE_B200_UNROLLED_NODES'] = '0' if mode == 'aten' else '1'
    torch.manual_seed(0)
    net = models.ODENet(3, n_filters=64, downsample='residual', tol=1e-3, adjoint=(mode == 'adjoint'), dropout=0.5).train().to(dev)
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    x = torch.rand(128, 3, 32, 32, device=dev); y = torch.randint(0, 10, (128,), device=dev)
    def step():
        opt.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(net(x), y).backward()
        opt.step()
    for _ in range(3): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): step()
    torch.cuda.synchronize()
    print(mode, 'train step b128: %.2f ms' % ((time.perf_counter() - t0) / 10 * 1e3), flush=True)
PY
timeout 200 python tools/pgd_latency.py 2>&1 | grep "^1 1\|^128 1"
