"""Evidence for the adjoint parity metric: how much do the REFERENCE's adjoint gradients
move when y(t1) is perturbed at float32 noise level?  Runs the oracle (pinned bit-for-bit to the
reference) on the committed golden case.  Result recorded in DESIGN.md:

    eps 1e-06: max-rel 2e-06 .. 5e-06      (no ReLU mask flips)
    eps 3e-06: max-rel 3e-03 .. 8e-03, L2-rel 1e-03 .. 3e-03
    eps 1e-05: max-rel 3e-02 .. 2e-01, L2-rel 5e-03 .. 1e-02
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dopri5_port, odefunc_port  # noqa: E402

g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'adjoint_cifar_n4.npz')))
p = {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('p.')}
params = [p[k] for k in odefunc_port.PARAM_ORDER]
t, out, go = torch.from_numpy(g['t']), torch.from_numpy(g['out']), torch.from_numpy(g['grad_out'])
f = lambda a, b: odefunc_port.odefunc_forward(p, a, b)
vj = lambda a, b, c: odefunc_port.odefunc_vjp(p, a, b, c)
gy, gt, gp = dopri5_port.adjoint_backward(f, params, t, out, go, 1e-3, 1e-3, vjp=vj)
for eps in (1e-6, 3e-6, 1e-5):
    for seed in (0, 1):
        torch.manual_seed(seed)
        o2 = out.clone()
        o2[-1] = o2[-1] * (1 + eps * torch.randn_like(o2[-1]))
        gy2, gt2, gp2 = dopri5_port.adjoint_backward(f, params, t, o2, go, 1e-3, 1e-3, vjp=vj)
        d = gy2 - gy
        per_img = [float(d[i].abs().max() / gy.abs().max()) for i in range(d.shape[0])]
        print('eps %.0e seed %d: grad_y0 max-rel %.2e l2-rel %.2e per-image %s | grad_params max-rel %.2e l2-rel %.2e' % (
            eps, seed, float(d.abs().max() / gy.abs().max()), float(d.norm() / gy.norm()), ['%.1e' % v for v in per_img],
            float((gp2 - gp).abs().max() / gp.abs().max()), float((gp2 - gp).norm() / gp.norm())))
