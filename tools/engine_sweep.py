"""Tuning aid: ODE-block forward time of the two 8x8 step engines over the batch size (which one serves small batches)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'neural-ode-features_b200'), ROOT]
import torch
from node_b200 import models
torch.manual_seed(0)
net = models.ODENet(3, n_filters=64, downsample='residual', tol=1e-3).eval().cuda()
for B in (16, 64, 128, 256, 384, 512, 592, 768, 1024, 1184, 2368):
    x = torch.rand(B, 3, 32, 32, device='cuda')
    row = []
    with torch.no_grad():
        h0 = net.downsample(x)
        for eng in ('1', '0', 'auto'):
            if eng == 'auto':
                os.environ.pop('NODE_B200_STEP8', None)
            else:
                os.environ['NODE_B200_STEP8'] = eng
            for _ in range(3):
                net.odeblock(h0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                net.odeblock(h0)
            torch.cuda.synchronize()
            row.append(1e3 * (time.perf_counter() - t0) / 20)
    print('batch %5d  dense %.3f ms  strip %.3f ms  auto %.3f ms' % (B, row[0], row[1], row[2]))
