#!/bin/bash
cat > /tmp/ab.py <<'PY'
import os, sys, time
sys.path.insert(0, '.')
import __graft_entry__ as ge
ge.build()
import torch
from node_b200 import models, solver, native
dev = 'cuda:0'
torch.zeros(1, device=dev)
t0 = time.perf_counter()
for _ in range(20000): native.stream_ptr()
print('stream_ptr us/call', (time.perf_counter() - t0) / 20000 * 1e6, 'slow' if native._SLOW_STREAM else 'fast', flush=True)
for mode in ('adjoint', 'nodes'):
    torch.manual_seed(0)
    net = models.ODENet(3, n_filters=64, downsample='residual', tol=1e-3, adjoint=(mode == 'adjoint'), dropout=0.5).train().to(dev)
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    x = torch.rand(128, 3, 32, 32, device=dev); y = torch.randint(0, 10, (128,), device=dev)
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
    print(mode, 'train step b128: %.2f ms (best of 3)' % best, flush=True)
PY
for m in 0 1 0 1; do NODE_B200_SLOW_STREAM=$m timeout 200 python /tmp/ab.py 2>&1 | tail -3; done
