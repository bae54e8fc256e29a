import os, sys
sys.path.insert(0, 'neural-ode-features_b200'); sys.path.insert(0, '.')
import torch, torch.nn as nn, torch.nn.functional as F
import __graft_entry__ as e; e.build()
from node_b200 import models, caller_grad, caller_ops
torch.backends.cudnn.allow_tf32 = False
DEV = 'cuda'
def rel(a, b): return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 150
torch.manual_seed(0)
for hw in (15, 8):
    blk = models.ResBlock(64, 64, stride=2, downsample=nn.Conv2d(64, 64, 1, 2, bias=False)).to(DEV)
    c = torch.randn(N, 64, hw, hw, device=DEV, requires_grad=True)
    sc = torch.randn(N, 64, hw, hw, device=DEV, requires_grad=True)
    gr = torch.randn(N, 64, hw, hw, device=DEV) * 1e-3
    ps = [blk.norm2.weight, blk.norm2.bias, blk.conv2.weight]
    ref = torch.autograd.grad(blk.conv2(torch.relu(blk.norm2(c))) + sc, [c, sc] + ps, gr)
    out = caller_ops.res_conv(blk.norm2, blk.conv2, c, sc)
    got = torch.autograd.grad(out, [c, sc] + ps, gr)
    print('tail', hw, [('%.1e' % rel(g, r)) for g, r in zip(got, ref)])
    hi = 2 * hw if hw == 15 else 15
    a = torch.relu(torch.randn(N, 64, hi, hi, device=DEV)).requires_grad_(True)
    gc = torch.randn(N, 64, hw, hw, device=DEV) * 1e-3
    gsc = torch.randn(N, 64, hw, hw, device=DEV) * 1e-3
    ps = [blk.conv1.weight, blk.downsample.weight]
    ref = torch.autograd.grad([blk.conv1(a), blk.downsample(a)], [a] + ps, [gc, gsc])
    o = caller_ops.res_head(blk.norm1, blk.conv1, blk.downsample, a)
    got = torch.autograd.grad(list(o), [a] + ps, [gc, gsc])
    print('head', hi, [('%.1e' % rel(g, r)) for g, r in zip(got, ref)], type(o[0].grad_fn).__name__)
conv0 = nn.Conv2d(3, 64, 3, 1).to(DEV); norm = nn.GroupNorm(32, 64).to(DEV)
x = torch.rand(N, 3, 32, 32, device=DEV)
go = torch.randn(N, 64, 30, 30, device=DEV) * 1e-3
ps = [conv0.weight, conv0.bias, norm.weight, norm.bias]
ref = torch.autograd.grad(torch.relu(norm(conv0(x))), ps, go)
got = torch.autograd.grad(caller_ops.stem_gn_relu(conv0, norm, x), ps, go)
print('stem', [('%.1e' % rel(g, r)) for g, r in zip(got, ref)])
c64 = nn.Conv2d(3, 64, 3, 1).to(DEV).double(); n64 = nn.GroupNorm(32, 64).to(DEV).double()
c64.load_state_dict({k: v.double() for k, v in conv0.state_dict().items()}); n64.load_state_dict({k: v.double() for k, v in norm.state_dict().items()})
ps64 = [c64.weight, c64.bias, n64.weight, n64.bias]
ref64 = torch.autograd.grad(torch.relu(n64(c64(x.double()))), ps64, go.double())
print('stem native vs f64', [('%.1e' % rel(g.double(), r)) for g, r in zip(got, ref64)])
print('stem aten32 vs f64', [('%.1e' % rel(g.double(), r)) for g, r in zip(ref, ref64)])
