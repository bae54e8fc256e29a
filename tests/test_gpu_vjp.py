"""GPU: the native adjoint kernels (K7: fused VJP + weight-gradient GEMM + fold) against the oracle's
hand-derived vector-Jacobian products (oracle/odefunc_port.py:odefunc_vjp, itself pinned against
torch.autograd of the reference op sequence in tests/test_oracle.py), for every feature-map shape the
reference's downsamplers produce, both time directions, and a batch that spans several super-tiles per CTA."""
import numpy as np
import pytest
import torch

from conftest import load_odefunc, odefunc_params
from oracle import odefunc_port

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
C = 64
WSZ = C * (C + 1) * 9
BLK = 2 * C + WSZ + C
TOL_VJP = 5e-5          # bf16x3 gradient operands: 2^-16 per product, fp32 accumulation


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def pieces(flat):
    out = {}
    for b in range(2):
        o = b * BLK
        out['dgamma%d' % (b + 1)] = flat[o:o + C]
        out['dbeta%d' % (b + 1)] = flat[o + C:o + 2 * C]
        out['dW%d' % (b + 1)] = flat[o + 2 * C:o + 2 * C + WSZ]
        out['dbias%d' % (b + 1)] = flat[o + 2 * C + WSZ:o + BLK]
    out['dgamma3'] = flat[2 * BLK:2 * BLK + C]
    out['dbeta3'] = flat[2 * BLK + C:2 * BLK + 2 * C]
    return out


@pytest.mark.parametrize('tsign', [1.0, -1.0])
@pytest.mark.parametrize('name,rep,n,engine', [
    ('cifar_res_n8', 1, 8, '0'), ('mnist_res_n5', 1, 5, None), ('mnist_conv_n9', 1, 9, None), ('mnist_oneshot_n3', 1, 3, None),
    ('cifar_oneshot_n3', 1, 3, None), ('cifar_res_n8', 80, 640, '0'),
    # the dense pair engine (k_vjp8): default above 444 images; forced for the small / ragged cases (a pair with one active
    # slot, a last super-tile of one image, a peer CTA without work)
    ('cifar_res_n8', 80, 640, None), ('cifar_res_n8', 81, 645, None), ('cifar_res_n8', 1, 8, '1'), ('cifar_res_n8', 1, 5, '1'),
    ('cifar_res_n8', 3, 21, '1'), ('cifar_res_n8', 320, 2560, None)])
def test_native_vjp_matches_oracle(native_lib, golden, monkeypatch, name, rep, n, engine, tsign):
    from node_b200 import solver
    if engine is not None:
        monkeypatch.setenv('NODE_B200_VJP8', engine)
    else:
        monkeypatch.delenv('NODE_B200_VJP8', raising=False)
    g = golden(name)
    func = load_odefunc(g, DEV)
    p = odefunc_params(g)
    h0 = torch.from_numpy(g['h0'])
    if rep > 1:                                     # 640 images: > 148 super-tiles, ragged tail (640 = 213*3 + 1)
        h0 = torch.cat([h0 * (1 + 0.003 * i) for i in range(rep)], 0)
    h0 = h0[:n].contiguous()
    gen = torch.Generator().manual_seed(7)
    adj = torch.randn(h0.shape, generator=gen) * 1e-2
    t = 0.37
    f, vy, vt, vp = solver.odefunc_vjp(func, t, h0.to(DEV), adj.to(DEV), tsign=tsign)
    torch.cuda.synchronize()
    rf, ry, rt, rp = odefunc_port.odefunc_vjp(p, torch.tensor(tsign * t), h0, -adj)     # adjoint.py:43 cotangent
    rf, ry, rt, rp = tsign * rf, tsign * ry, tsign * rt, tsign * rp                      # misc.py:184-187 wrapper
    assert rel(f.cpu(), rf) < 2e-5
    assert rel(vy.cpu(), ry) < TOL_VJP
    assert abs(float(vt) - float(rt)) <= TOL_VJP * abs(float(rt)) + 1e-9
    got, ref = pieces(vp.cpu()), pieces(rp)
    for k in ref:
        assert rel(got[k], ref[k]) < TOL_VJP, k
    assert vp.shape == (solver.N_PARAMS_64,) == rp.shape


def test_native_vjp_is_deterministic(native_lib, golden):
    """Partials are folded in a fixed order: two evaluations give bit-identical gradients."""
    from node_b200 import solver
    g = golden('cifar_res_n8')
    func = load_odefunc(g, DEV)
    h0 = torch.cat([torch.from_numpy(g['h0'])] * 40, 0).to(DEV)
    adj = torch.randn(h0.shape, generator=torch.Generator().manual_seed(3)).to(DEV)
    a = [o.clone() for o in solver.odefunc_vjp(func, 0.1, h0, adj)]
    b = solver.odefunc_vjp(func, 0.1, h0, adj)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_adjoint_backward_runs_on_the_native_vjp(native_lib, golden):
    """odeint_adjoint backward (adjoint.py:23-102): identical NFE-B / step sequence, no autograd inside the solve."""
    from node_b200 import odeint_adjoint, solver
    g = golden('adjoint_cifar_n4')
    func = load_odefunc(g, DEV).train()
    h0 = torch.from_numpy(g['h0']).to(DEV).requires_grad_(True)
    t = torch.from_numpy(g['t']).to(DEV)
    out = odeint_adjoint(func, h0, t, rtol=1e-3, atol=1e-3, method='dopri5')
    func.nfe = 0
    out.backward(torch.from_numpy(g['grad_out']).to(DEV))
    assert solver.last_stats['adjoint_vjp'] == 'native'
    assert func.nfe == int(g['nfe_b'])
    assert solver.last_stats['route'] == 'generic' and solver.last_stats['status'] == 0
