"""Small invocations of the kernels added late in round 2, for compute-sanitizer (memcheck / racecheck / synccheck):
the adjoint interval as a CUDA graph WHILE loop (node_b200_adjoint_solve), the wide augmented dynamics (node_b200_wide_vjp: wide8
raw operand + implicit GEMM with the output-channel split, GroupNorm backward, batch column sums, block weight gradients) and the
unrolled route (native dynamics + VJP under autograd)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'neural-ode-features_b200'), ROOT]
import torch
import __graft_entry__ as entry
entry.build()
from node_b200 import models, odeint, odeint_adjoint, solver, wide

dev = 'cuda:0'
torch.manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
t = torch.tensor([0.0, 1.0], device=dev)
if which in ('all', 'adjoint'):
    f = models.ODEfunc(64).to(dev)
    h = torch.randn(5, 64, 8, 8, device=dev).requires_grad_(True)
    out = odeint_adjoint(f, h, t, rtol=1e-2, atol=1e-2, method='dopri5')
    out.backward(torch.randn_like(out))
    torch.cuda.synchronize()
    print('adjoint (device loop):', solver.last_stats.get('adjoint_loop'), 'finite', bool(torch.isfinite(h.grad).all()), flush=True)
if which in ('all', 'wide'):
    for C, n in ((128, 3), (256, 9)):
        f = models.ODEfunc(C).to(dev)
        y = torch.randn(n, C, 8, 8, device=dev)
        a = torch.randn(n, C, 8, 8, device=dev)
        P = sum(p.numel() for p in f.parameters())
        dst = (torch.empty_like(y), torch.empty_like(y), torch.empty((), device=dev), torch.empty(P, device=dev))
        with torch.no_grad():
            wide.WideAugmented(f).eval_into(torch.tensor(0.3, device=dev), (y, a), dst, 1.0)
        torch.cuda.synchronize()
        print('wide vjp C', C, 'batch', n, 'finite', bool(torch.isfinite(dst[1]).all() and torch.isfinite(dst[3]).all()), flush=True)
    f = models.ODEfunc(128).to(dev)
    y = torch.randn(2, 128, 7, 7, device=dev)
    a = torch.randn(2, 128, 7, 7, device=dev)
    P = sum(p.numel() for p in f.parameters())
    dst = (torch.empty_like(y), torch.empty_like(y), torch.empty((), device=dev), torch.empty(P, device=dev))
    with torch.no_grad():
        wide.WideAugmented(f).eval_into(torch.tensor(0.3, device=dev), (y, a), dst, -1.0)
    torch.cuda.synchronize()
    print('wide vjp 7x7 (block path) finite', bool(torch.isfinite(dst[1]).all()), flush=True)
if which in ('all', 'unrolled'):
    f = models.ODEfunc(64).to(dev)
    h = torch.randn(3, 64, 8, 8, device=dev).requires_grad_(True)
    out = odeint(f, h, t, rtol=1e-2, atol=1e-2, method='dopri5')
    out.backward(torch.randn_like(out))
    torch.cuda.synchronize()
    print('unrolled:', solver.last_stats.get('route'), 'finite', bool(torch.isfinite(h.grad).all()), flush=True)
print('done')
