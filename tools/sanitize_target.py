"""Small invocations of every tensor-core kernel of the path, for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool racecheck python tools/sanitize_target.py
k_step (strip engine), k_step8 (dense pair engine), k_vjp / k_vjp8, k_wgrad, k_convs2, k_resconv, stem, head, RK kernels."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'neural-ode-features_b200'), ROOT]
import torch
import __graft_entry__ as entry
entry.build()
from node_b200 import models, solver

torch.manual_seed(0)
dev = torch.device('cuda', 0)
net = models.ODENet(3, n_filters=64, downsample='residual', tol=1e-3).eval().to(dev)
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
engines = sys.argv[2] if len(sys.argv) > 2 else '01'       # '0': strip engines only, '1': dense pair (cluster) engines only
with torch.no_grad():
    if which in ('all', 'forward'):
        for dense, n in (('0', 7), ('1', 20)):            # strip engine / dense pair engine (ragged: 20 = 5 super-tiles)
            if dense not in engines:
                continue
            os.environ['NODE_B200_STEP8'] = dense
            out = net(torch.rand(n, 3, 32, 32, device=dev))
            torch.cuda.synchronize()
            print('forward engine', dense, 'batch', n, 'nfe', net.odeblock.odefunc.nfe, 'finite', bool(torch.isfinite(out).all()), flush=True)
    if which in ('all', 'vjp'):
        func = net.odeblock.odefunc
        for dense, n in (('0', 7), ('1', 20)):
            if dense not in engines:
                continue
            os.environ['NODE_B200_VJP8'] = dense
            y = torch.randn(n, 64, 8, 8, device=dev)
            a = torch.randn(n, 64, 8, 8, device=dev) * 1e-2
            f, vy, vt, vp = solver.odefunc_vjp(func, 0.3, y, a)
            torch.cuda.synchronize()
            print('vjp engine', dense, 'batch', n, 'finite', bool(torch.isfinite(vy).all() and torch.isfinite(vp).all()), flush=True)
print('done')
