"""Stage-by-stage check of the wide augmented dynamics against float64 autograd (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.build()
import copy
import torch
import torch.nn.functional as F
from node_b200 import models, wide
DEV = 'cuda:0'
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
C, hw = int(os.environ.get('C', 256)), 8
for n in [int(v) for v in os.environ.get('NS', '5,32,64,70').split(',')]:
    torch.manual_seed(C + hw + n)
    f = models.ODEfunc(C).to(DEV)
    with torch.no_grad():
        for norm in (f.norm1, f.norm2, f.norm3):
            norm.weight.uniform_(0.5, 1.5)
            norm.bias.uniform_(-0.5, 0.5)
    y = torch.randn(n, C, hw, hw, device=DEV) * 1.5 + 0.2
    adj = torch.randn(n, C, hw, hw, device=DEV)
    t = 0.37
    P = sum(p.numel() for p in f.parameters())
    dst = (torch.empty_like(y), torch.empty_like(y), torch.empty((), device=DEV), torch.empty(P, device=DEV))
    aug = wide.WideAugmented(f)
    with torch.no_grad():
        aug.eval_into(torch.tensor(t, device=DEV), (y, adj), dst, 1.0)
    v = aug.dyn._v
    f64 = copy.deepcopy(f).double()
    tt = torch.tensor(t, dtype=torch.float64, device=DEV)
    yy = y.double().requires_grad_(True)
    a1 = F.relu(f64.norm1(yy)); a1.retain_grad()
    c1 = f64.conv1(tt, a1); c1.retain_grad()
    a2 = F.relu(f64.norm2(c1)); a2.retain_grad()
    c2 = f64.conv2(tt, a2); c2.retain_grad()
    out = f64.norm3(c2)
    out.backward(-adj.double())
    def e(got, want):
        return '%.1e/%.1e' % (float((got.double() - want).abs().max() / want.abs().max()), float((got.double() - want).norm() / want.norm()))
    b1 = f.conv1._layer.bias.view(1, C, 1, 1); b2 = f.conv2._layer.bias.view(1, C, 1, 1)
    tm1 = aug.dyn._tmap[0].view(1, C, hw, hw); tm2 = aug.dyn._tmap[1].view(1, C, hw, hw)
    print('n', n, 'a1', e(v['a1'], a1), 'c1', e(v['c1'] + b1 + t * tm1, c1), 'a2', e(v['a2'], a2), 'c2', e(v['c2'] + b2 + t * tm2, c2), 'f', e(dst[0], out),
          'gc2', e(v['gc2'], c2.grad), 'ga2', 'n/a', 'gc1', e(v['gc1'], c1.grad), 'gr1(a1.grad)', e(v['gr'], a1.grad), 'vy', e(dst[1], yy.grad), flush=True)
