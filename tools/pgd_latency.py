"""Batch-1 PGD iteration latency (cfg5) with the adjoint interval as one C call vs one controller read per attempted step."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.build()
import torch
from node_b200 import models, solver

dev = 'cuda:0'
torch.manual_seed(0)
net = models.ODENet(3, n_filters=64, downsample='residual', tol=1e-3, adjoint=True).eval().to(dev)
out = {}
for batch in (1, 128):
    x0 = torch.rand(batch, 3, 32, 32, device=dev)
    lab = torch.randint(0, 10, (batch,), device=dev)
    for mode in ('1', '0'):
        os.environ['NODE_B200_ADJOINT_SOLVE'] = mode
        for tol in (1e-3, 1e-1):
            net.odeblock.tol = tol
            def it():
                x = x0.clone().requires_grad_(True)
                loss = torch.nn.functional.cross_entropy(net(x), lab)
                g, = torch.autograd.grad(loss, x)
                return g
            for _ in range(3):
                it()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                it()
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) / 20 * 1e3
            nfe = net.nfe(reset=True)
            # the backward alone
            x = x0.clone().requires_grad_(True)
            loss = torch.nn.functional.cross_entropy(net(x), lab)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            torch.autograd.grad(loss, x)
            torch.cuda.synchronize()
            bms = (time.perf_counter() - t0) * 1e3
            out['b%d_onecall%s_tol%g' % (batch, mode, tol)] = dict(ms_per_iteration=ms, backward_ms=bms, loop=solver.last_stats.get('adjoint_loop'))
            print(batch, mode, tol, ms, bms, flush=True)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/r02ae_pgd_latency.json', 'w'), indent=1)
