"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement of the ODE-Net dynamics function of fabiocarrara/neural-ode-features
(/root/reference/model.py:313-348): GroupNorm -> ReLU -> ConcatConv2d(t) -> GroupNorm ->
ReLU -> ConcatConv2d(t) -> GroupNorm, with GroupNorm(min(32, C), C), eps 1e-5
(model.py:268-271) and ConcatConv2d = conv2d(cat([t*ones, x], 1), W[C, C+1, 3, 3], b, pad 1)
(model.py:313-323).

Three views of the same function:
  odefunc_forward         the reference op sequence (cat + conv2d + group_norm);
  odefunc_forward_folded  the algebra the CUDA kernels use: the time plane folded into a
                          position dependent bias  t * Tmap[c, h, w]  (SURVEY fact 3);
  odefunc_vjp             hand-derived vector-Jacobian products (the arithmetic the adjoint
                          kernels implement), pinned against torch.autograd in
                          tests/test_oracle.py.

Parity status: PINNED (tools/make_golden.py checks odefunc_forward bit-for-bit against the
imported reference module with the same state_dict).

Parameters travel as a dict with the reference's state_dict key names relative to
`odeblock.odefunc.`:  norm{1,2,3}.{weight,bias}, conv{1,2}._layer.{weight,bias}.
"""
import torch
import torch.nn.functional as F

PARAM_ORDER = (
    'norm1.weight', 'norm1.bias', 'conv1._layer.weight', 'conv1._layer.bias',
    'norm2.weight', 'norm2.bias', 'conv2._layer.weight', 'conv2._layer.bias',
    'norm3.weight', 'norm3.bias',
)
EPS = 1e-5


def n_groups(C):
    return min(32, C)


def params_from_module(odefunc):
    sd = dict(odefunc.named_parameters())
    return {k: sd[k] for k in PARAM_ORDER}


def concat_conv(t, x, W, b):
    """model.py:320-323."""
    tt = torch.ones_like(x[:, :1, :, :]) * t
    return F.conv2d(torch.cat([tt, x], 1), W, b, stride=1, padding=1)


def odefunc_forward(p, t, x):
    """model.py:339-348 (ReLU out of place: same values)."""
    C = x.shape[1]
    G = n_groups(C)
    out = F.group_norm(x, G, p['norm1.weight'], p['norm1.bias'], EPS)
    out = torch.relu(out)
    out = concat_conv(t, out, p['conv1._layer.weight'], p['conv1._layer.bias'])
    out = F.group_norm(out, G, p['norm2.weight'], p['norm2.bias'], EPS)
    out = torch.relu(out)
    out = concat_conv(t, out, p['conv2._layer.weight'], p['conv2._layer.bias'])
    out = F.group_norm(out, G, p['norm3.weight'], p['norm3.bias'], EPS)
    return out


def time_map(W, H, Wd):
    """Tmap[c,h,w] = conv2d(ones(1,1,H,W), W[:, 0:1], padding=1): 9 distinct values per c."""
    ones = torch.ones(1, 1, H, Wd, dtype=W.dtype, device=W.device)
    return F.conv2d(ones, W[:, 0:1], None, stride=1, padding=1)[0]


def odefunc_forward_folded(p, t, x):
    C, H, Wd = x.shape[1:]
    G = n_groups(C)
    out = torch.relu(F.group_norm(x, G, p['norm1.weight'], p['norm1.bias'], EPS))
    W1, W2 = p['conv1._layer.weight'], p['conv2._layer.weight']
    out = F.conv2d(out, W1[:, 1:], None, padding=1) + (p['conv1._layer.bias'][:, None, None] + t * time_map(W1, H, Wd))
    out = torch.relu(F.group_norm(out, G, p['norm2.weight'], p['norm2.bias'], EPS))
    out = F.conv2d(out, W2[:, 1:], None, padding=1) + (p['conv2._layer.bias'][:, None, None] + t * time_map(W2, H, Wd))
    return F.group_norm(out, G, p['norm3.weight'], p['norm3.bias'], EPS)


# ---- hand-derived backward -----------------------------------------------------------------

def _gn_fwd(x, G, gamma, beta):
    N, C, H, Wd = x.shape
    xg = x.reshape(N, G, -1)
    mean = xg.mean(-1, keepdim=True)
    var = xg.var(-1, unbiased=False, keepdim=True)
    rstd = 1.0 / torch.sqrt(var + EPS)
    xhat = ((xg - mean) * rstd).reshape(N, C, H, Wd)
    return xhat * gamma[None, :, None, None] + beta[None, :, None, None], xhat, rstd


def _gn_bwd(g, xhat, rstd, G, gamma):
    """Returns (dx, dgamma, dbeta) for y = xhat*gamma + beta, stats per (n, group)."""
    N, C, H, Wd = g.shape
    dgamma = (g * xhat).sum((0, 2, 3))
    dbeta = g.sum((0, 2, 3))
    gg = (g * gamma[None, :, None, None]).reshape(N, G, -1)
    xh = xhat.reshape(N, G, -1)
    m1 = gg.mean(-1, keepdim=True)
    m2 = (gg * xh).mean(-1, keepdim=True)
    dx = (rstd * (gg - m1 - xh * m2)).reshape(N, C, H, Wd)
    return dx, dgamma, dbeta


def _conv_bwd(g, a, t, W):
    """Backward of concat_conv wrt (input a, t, W, b).  g: dL/dout [N,C,H,W]."""
    N, C, H, Wd = a.shape
    Wx = W[:, 1:]
    da = F.conv_transpose2d(g, Wx, None, stride=1, padding=1)
    db = g.sum((0, 2, 3))
    dt = (g * time_map(W, H, Wd)[None]).sum()
    # weight gradient: dW[o, i, dy, dx] = sum_{n,h,w} g[n,o,h,w] * in[n,i,h+dy-1,w+dx-1]
    ap = F.pad(torch.cat([torch.ones_like(a[:, :1]) * t, a], 1), (1, 1, 1, 1))
    dW = torch.empty_like(W)
    for dy in range(3):
        for dx in range(3):
            win = ap[:, :, dy:dy + H, dx:dx + Wd]
            dW[:, :, dy, dx] = torch.einsum('nohw,nihw->oi', g, win)
    return da, dt, dW, db


def odefunc_vjp(p, t, x, a):
    """(f, vjp_x, vjp_t, flat vjp_params) with cotangent `a` on f; params in PARAM_ORDER."""
    C = x.shape[1]
    G = n_groups(C)
    n1, xh1, r1 = _gn_fwd(x, G, p['norm1.weight'], p['norm1.bias'])
    a1 = torch.relu(n1)
    c1 = concat_conv(t, a1, p['conv1._layer.weight'], p['conv1._layer.bias'])
    n2, xh2, r2 = _gn_fwd(c1, G, p['norm2.weight'], p['norm2.bias'])
    a2 = torch.relu(n2)
    c2 = concat_conv(t, a2, p['conv2._layer.weight'], p['conv2._layer.bias'])
    f, xh3, r3 = _gn_fwd(c2, G, p['norm3.weight'], p['norm3.bias'])

    dc2, dg3, db3 = _gn_bwd(a, xh3, r3, G, p['norm3.weight'])
    da2, dt2, dW2, dbias2 = _conv_bwd(dc2, a2, t, p['conv2._layer.weight'])
    dn2 = da2 * (n2 > 0).to(da2.dtype)
    dc1, dg2, db2 = _gn_bwd(dn2, xh2, r2, G, p['norm2.weight'])
    da1, dt1, dW1, dbias1 = _conv_bwd(dc1, a1, t, p['conv1._layer.weight'])
    dn1 = da1 * (n1 > 0).to(da1.dtype)
    dx, dg1, db1 = _gn_bwd(dn1, xh1, r1, G, p['norm1.weight'])
    flat = torch.cat([v.reshape(-1) for v in (dg1, db1, dW1, dbias1, dg2, db2, dW2, dbias2, dg3, db3)])
    return f, dx, (dt1 + dt2).reshape(()), flat
