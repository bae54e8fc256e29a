"""ODE block forward + odeint_adjoint backward at n_filters = 256 (the paper's CIFAR setting): native wide augmented dynamics vs the
autograd / cuDNN route, and the launch count of one backward."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.build()
import torch
from node_b200 import models, odeint_adjoint, solver
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = 'cuda:0'
out = {}
for C, N in ((256, 128), (256, 1024), (128, 128)):
    torch.manual_seed(0)
    f = models.ODEfunc(C).to(dev)
    h0 = (torch.randn(N, C, 8, 8, device=dev) * 0.5)
    t = torch.tensor([0.0, 1.0], device=dev)
    go = torch.randn(N, C, 8, 8, device=dev) * 1e-3
    for mode in ('1', '0'):
        os.environ['NODE_B200_NATIVE_VJP'] = mode
        def step():
            h = h0.clone().requires_grad_(True)
            for p in f.parameters():
                p.grad = None
            f.nfe = 0
            o = odeint_adjoint(f, h, t, rtol=1e-3, atol=1e-3, method='dopri5')[-1]
            nf = f.nfe
            o.backward(go)
            return nf, f.nfe - nf
        for _ in range(2):
            nfe = step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            step()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / reps * 1e3
        out['C%d_N%d_%s' % (C, N, 'native' if mode == '1' else 'autograd')] = dict(ms=ms, nfe_f=nfe[0], nfe_b=nfe[1], vjp=solver.last_stats.get('adjoint_vjp'))
        print(C, N, mode, '%.1f ms' % ms, nfe, solver.last_stats.get('adjoint_vjp'), flush=True)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/r02ak_wide_train.json', 'w'), indent=1)
