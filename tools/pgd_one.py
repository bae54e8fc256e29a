"""One batch-1 PGD iteration (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.build()
import torch
from node_b200 import models
dev = 'cuda:0'
torch.manual_seed(0)
B = int(os.environ.get('PGD_BATCH', '1'))
net = models.ODENet(3, n_filters=64, downsample='residual', tol=1e-3, adjoint=True).eval().to(dev)
x0 = torch.rand(B, 3, 32, 32, device=dev)
lab = torch.randint(0, 10, (B,), device=dev)
for i in range(3):
    x = x0.clone().requires_grad_(True)
    loss = torch.nn.functional.cross_entropy(net(x), lab)
    torch.cuda.synchronize()
    if i == 2:
        torch.cuda.cudart().cudaProfilerStart()
    g, = torch.autograd.grad(loss, x)
    torch.cuda.synchronize()
    if i == 2:
        torch.cuda.cudart().cudaProfilerStop()
