#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_unrolled.py -x -q -k native_combination 2>&1 | grep -E "assert|Error|passed|failed" | head -8
timeout 300 python - <<'PY'
import os, sys, time, cProfile, pstats
sys.path.insert(0, '.')
import __graft_entry__ as ge
ge.build()
import torch
from node_b200 import models, solver
dev = 'cuda:0'
torch.manual_seed(0)
net = models.ODENet(3, n_filters=64, downsample='residual', tol=1e-3, adjoint=False, dropout=0.5).train().to(dev)
opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
x = torch.rand(128, 3, 32, 32, device=dev); y = torch.randint(0, 10, (128,), device=dev)
def step():
    opt.zero_grad(set_to_none=True)
    torch.nn.functional.cross_entropy(net(x), y).backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats('cumulative').print_stats(28)
PY
