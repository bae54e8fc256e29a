"""Bring-up aid for the adjoint kernels (K7): per-component errors of the native VJP against the oracle,
for every golden shape.  Run on the GPU box."""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'neural-ode-features_b200'), ROOT, os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import __graft_entry__ as entry
entry.build()
from conftest import load_odefunc, odefunc_params
from node_b200 import solver, native
from oracle import odefunc_port

DEV = 'cuda:0'


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run(name, tsign, N_rep=1):
    g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz')))
    func = load_odefunc(g, DEV)
    p = odefunc_params(g)
    h0 = torch.from_numpy(g['h0'])
    if N_rep > 1:
        h0 = torch.cat([h0 * (1 + 0.01 * i) for i in range(N_rep)], 0)
    gen = torch.Generator().manual_seed(7)
    adj = torch.randn(h0.shape, generator=gen) * 1e-2
    t = 0.37
    f, vy, vt, vp = solver.odefunc_vjp(func, t, h0.to(DEV), adj.to(DEV), tsign=tsign)
    torch.cuda.synchronize()
    rf, ry, rt, rp = odefunc_port.odefunc_vjp(p, torch.tensor(tsign * t), h0, -adj)
    rf, ry, rt, rp = tsign * rf, tsign * ry, tsign * rt, tsign * rp
    C = 64
    Wsz = C * (C + 1) * 9
    blk = 2 * C + Wsz + C
    vp, f, vy, vt = vp.cpu(), f.cpu(), vy.cpu(), vt.cpu()
    out = dict(f=rel(f, rf), vjp_y=rel(vy, ry), vjp_t=abs(float(vt) - float(rt)) / max(abs(float(rt)), 1e-30))
    for b in range(2):
        o = b * blk
        out['dgamma%d' % (b + 1)] = rel(vp[o:o + C], rp[o:o + C])
        out['dbeta%d' % (b + 1)] = rel(vp[o + C:o + 2 * C], rp[o + C:o + 2 * C])
        Wg, Wr = vp[o + 2 * C:o + 2 * C + Wsz].view(C, C + 1, 9), rp[o + 2 * C:o + 2 * C + Wsz].view(C, C + 1, 9)
        out['dW%d' % (b + 1)] = rel(Wg[:, 1:], Wr[:, 1:])
        out['dWt%d' % (b + 1)] = rel(Wg[:, 0], Wr[:, 0])
        out['dbias%d' % (b + 1)] = rel(vp[o + 2 * C + Wsz:o + blk], rp[o + 2 * C + Wsz:o + blk])
    o = 2 * blk
    out['dgamma3'] = rel(vp[o:o + C], rp[o:o + C])
    out['dbeta3'] = rel(vp[o + C:o + 2 * C], rp[o + C:o + 2 * C])
    # the GEMM alone, against torch applied to the GPU's own operands
    N, _, H, W = h0.shape
    vws = solver._vjp_workspace(torch.device(DEV), N, C, H, W)
    L = native.lib()
    def buf(which):
        ptr_ = L.node_b200_vjp_buffer(native.ptr(vws), which, N, C, H, W)
        off = ptr_ - vws.data_ptr()
        return vws[off:off + N * C * H * W * 4].view(torch.float32).view(N, C, H, W).cpu()
    R1, R2, G1, G2 = buf(0), buf(1), buf(2), buf(3)
    for b, (R, G) in enumerate(((R1, G1), (R2, G2))):
        ap = torch.nn.functional.pad(R.double(), (1, 1, 1, 1))
        dW = torch.empty(C, C, 9, dtype=torch.float64)
        for dy in range(3):
            for dx in range(3):
                dW[:, :, dy * 3 + dx] = torch.einsum('nohw,nihw->oi', G.double(), ap[:, :, dy:dy + H, dx:dx + W])
        o = b * blk
        Wg = vp[o + 2 * C:o + 2 * C + Wsz].view(C, C + 1, 9)
        out['gemm%d' % (b + 1)] = rel(Wg[:, 1:].double(), tsign * dW)
    print('%-20s s=%+d N=%d ' % (name, tsign, h0.shape[0]) + ' '.join('%s=%.1e' % kv for kv in out.items()), flush=True)


if __name__ == '__main__':
    names = ['cifar_res_n8', 'mnist_res_n5', 'mnist_conv_n9', 'mnist_oneshot_n3', 'cifar_oneshot_n3']
    for nm in names:
        run(nm, 1)
    run('cifar_res_n8', -1)
    run('cifar_res_n8', 1, N_rep=64)      # 512 images: several super-tiles per CTA
