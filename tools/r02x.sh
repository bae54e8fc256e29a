#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
import os, sys, time
sys.path.insert(0, 'neural-ode-features_b200'); sys.path.insert(0, '.')
import torch, torch.nn.functional as F
import __graft_entry__ as e; e.build()
from node_b200 import models
torch.backends.cudnn.allow_tf32 = False
dev = 'cuda'
for B in (1, 128):
    for mode in ('0', '1', '0', '1'):
        os.environ['NODE_B200_ADJOINT_STEP'] = mode
        torch.manual_seed(0)
        net = models.ODENet(3, n_filters=64, downsample='residual', tol=1e-3, adjoint=True, dropout=0.5).to(dev).train()
        opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
        x = torch.rand(B, 3, 32, 32, device=dev); y = torch.randint(0, 10, (B,), device=dev)
        def step():
            opt.zero_grad(set_to_none=True)
            loss = F.cross_entropy(net(x), y); loss.backward(); opt.step()
        for _ in range(5): step()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): step()
        torch.cuda.synchronize()
        print('batch %d adjoint_step=%s: %.2f ms per training step' % (B, mode, (time.perf_counter() - t0) / 20 * 1e3))
PY
