import os, sys, time
sys.path.insert(0, 'neural-ode-features_b200'); sys.path.insert(0, '.')
import torch
import __graft_entry__ as e; e.build()
from node_b200 import models, solver
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = 'cuda'
torch.manual_seed(0)
net = models.ODENet(3, n_filters=256, downsample='residual', tol=1e-3).eval().to(dev)
for B in (256, 1024, 2048, 2368):
    x = torch.rand(B, 3, 32, 32, device=dev)
    for mode in ('1', '0'):
        os.environ['NODE_B200_WIDE'] = mode
        with torch.no_grad():
            for _ in range(2): net(x)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0 = net.downsample(x)
            a.record()
            for _ in range(3): net.odeblock(h0)
            b.record(); torch.cuda.synchronize()
            ode_ms = a.elapsed_time(b) / 3
            a.record()
            for _ in range(3): net(x)
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 3
        print('B %d wide=%s route %s nfe %s: forward %.2f ms = %.0f img/s; ODE block %.2f ms (%.1f TFLOP/s algorithmic)' % (
            B, mode, solver.last_stats['route'], solver.last_stats['nfe'], ms, B / ms * 1e3, ode_ms, B * 26 * 2 * 2 * 9 * 256 * 256 * 64 / ode_ms / 1e9))
