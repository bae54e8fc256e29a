"""Per-evaluation latency of the adjoint's augmented dynamics (node_b200_odefunc_vjp) and of the forward dynamics at small batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.build()
import torch
from node_b200 import models, solver
dev = 'cuda:0'
torch.manual_seed(0)
f = models.ODEfunc(64).to(dev)
reps = int(os.environ.get('REPS', '50'))
for hw in (8,):
    for n in [int(v) for v in os.environ.get('BATCHES', '1,8,128').split(',')]:
        y = torch.randn(n, 64, hw, hw, device=dev)
        a = torch.randn(n, 64, hw, hw, device=dev)
        t = torch.tensor(0.3, device=dev)
        out = solver.odefunc_vjp(f, t, y, a)
        o2 = solver.odefunc_forward(f, 0.3, y)
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        for _ in range(reps):
            solver.odefunc_vjp(f, t, y, a, out=out)
        e1.record()
        for _ in range(reps):
            solver.odefunc_forward(f, 0.3, y)
        e2.record()
        torch.cuda.synchronize()
        print('hw %d batch %d: vjp eval %.1f us, forward eval %.1f us' % (hw, n, e0.elapsed_time(e1) / reps * 1e3, e1.elapsed_time(e2) / reps * 1e3), flush=True)
