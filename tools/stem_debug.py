import sys
sys.path.insert(0, 'neural-ode-features_b200'); sys.path.insert(0, '.')
import torch, torch.nn as nn
import __graft_entry__ as e; e.build()
from node_b200 import caller_ops
torch.backends.cudnn.allow_tf32 = False
DEV = 'cuda'
def rel(a, b): return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
for affine in (False, True):
  for n in (150, 296, 297, 450, 700, 1500):
    torch.manual_seed(8)
    conv0 = nn.Conv2d(3, 64, 3, 1).to(DEV); norm = nn.GroupNorm(32, 64).to(DEV)
    if affine:
        with torch.no_grad():
            norm.weight.uniform_(0.5, 1.5); norm.bias.uniform_(-0.5, 0.5)
    x = torch.rand(n, 3, 32, 32, device=DEV)
    go = torch.randn(n, 64, 30, 30, device=DEV) * 1e-3
    ps = [conv0.weight, conv0.bias, norm.weight, norm.bias]
    out = caller_ops.stem_gn_relu(conv0, norm, x)
    got = torch.autograd.grad(out, ps, go)
    c64, n64 = nn.Conv2d(3, 64, 3, 1).to(DEV).double(), nn.GroupNorm(32, 64).to(DEV).double()
    c64.load_state_dict({k: v.double() for k, v in conv0.state_dict().items()})
    n64.load_state_dict({k: v.double() for k, v in norm.state_dict().items()})
    o64 = torch.relu(n64(c64(x.double())))
    ref = torch.autograd.grad(o64, [c64.weight, c64.bias, n64.weight, n64.bias], go.double())
    o32 = torch.relu(norm(conv0(x)))
    flips_native = int(((out > 0) != (o64 > 0)).sum()); flips_aten = int(((o32 > 0) != (o64 > 0)).sum())
    print('affine', affine, 'n', n, ['%.1e' % rel(g.double(), r) for g, r in zip(got, ref)], 'mask flips native fwd', flips_native, 'aten', flips_aten)
