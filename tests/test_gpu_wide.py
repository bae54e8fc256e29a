"""GPU: the wide ODE-Net dynamics (n_filters = 128 / 256, reference reproduce.sh:21 and model.py:326-348) as 64-channel blocks on
the tcgen05 engine (node_b200/wide.py) against the modules' own ATen / cuDNN fp32 ops."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(autouse=True)
def _fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize('C,hw,n', [(256, 8, 5), (128, 8, 3), (256, 7, 4), (256, 8, 301)])
def test_wide_dynamics_match_module(native_lib, C, hw, n):
    from node_b200 import models, wide
    torch.manual_seed(C + hw)
    f = models.ODEfunc(C).to(DEV)
    with torch.no_grad():
        for norm in (f.norm1, f.norm2, f.norm3):
            norm.weight.uniform_(0.5, 1.5)
            norm.bias.uniform_(-0.5, 0.5)
        y = torch.randn(n, C, hw, hw, device=DEV) * 1.5 + 0.2
        for t in (0.0, 0.37, -0.8):
            ref = f(torch.tensor(t, device=DEV), y)
            got = wide.WideDynamics.of(f)(t, y)
            assert rel(got, ref) <= 3e-5, (t, rel(got, ref))
    assert f.nfe == 6


@pytest.mark.parametrize('C,n', [(256, 1), (256, 5), (256, 8), (128, 3), (256, 601), (128, 594), (256, 200), (256, 128), (128, 130)])
def test_wide8_conv_pieces(native_lib, C, n):
    """wide8 engine, piece by piece: operand image + implicit GEMM against GroupNorm -> ReLU -> cuDNN fp32 convolution."""
    import torch.nn.functional as F
    from node_b200 import models, wide, native
    torch.manual_seed(C + n)
    f = models.ODEfunc(C).to(DEV)
    with torch.no_grad():
        for norm in (f.norm1, f.norm2, f.norm3):
            norm.weight.uniform_(0.5, 1.5)
            norm.bias.uniform_(-0.5, 0.5)
        y = torch.randn(n, C, 8, 8, device=DEV) * 1.5 + 0.2
        wd = wide.WideDynamics.of(f)
        wd._prepare8()
        lib = native_lib
        op = torch.zeros(lib.node_b200_wide8_operand_bytes(n, C), dtype=torch.uint8, device=DEV)
        out = torch.empty_like(y)
        t = torch.tensor(0.4, device=DEV)
        for which, (norm, conv) in enumerate(((f.norm1, f.conv1._layer), (f.norm2, f.conv2._layer))):
            bias = f.conv1._layer.bias if which == 1 else None
            native.check(lib.node_b200_wide8_gn_operand(native.ptr(wd._ws8), which, native.ptr(y), native.ptr(op), native.ptr(norm.weight),
                                                        native.ptr(norm.bias), native.ptr(bias) if bias is not None else None, native.ptr(t),
                                                        1.0, n, C, native.stream_ptr()), 'gn_operand')
            native.check(lib.node_b200_wide8_conv(native.ptr(wd._ws8), which, native.ptr(op), native.ptr(out), n, C, native.stream_ptr()), 'conv')
            x = y
            if which == 1:
                tmap = F.conv2d(torch.ones(1, 1, 8, 8, device=DEV), f.conv1._layer.weight[:, :1], padding=1)
                x = y + bias.view(1, C, 1, 1) + t * tmap
            ref = F.conv2d(torch.relu(norm(x)), conv.weight[:, 1:], padding=1)
            assert rel(out, ref) <= 2e-5, (which, rel(out, ref))
        assert wd.watchdog() == 0


def test_wide8_matches_block_path(native_lib):
    from node_b200 import models, wide
    torch.manual_seed(9)
    f = models.ODEfunc(256).to(DEV)
    y = torch.randn(7, 256, 8, 8, device=DEV)
    with torch.no_grad():
        got = wide.WideDynamics.of(f)(0.3, y)
        os.environ['NODE_B200_WIDE8'] = '0'
        try:
            ref = wide.WideDynamics.of(f)(0.3, y)
        finally:
            os.environ.pop('NODE_B200_WIDE8', None)
    assert rel(got, ref) <= 2e-5, rel(got, ref)


@pytest.mark.parametrize('times', [[0.0, 1.0], [0.0, 0.3, 1.0], [1.0, 0.0]])
def test_wide_solve_matches_module_route(native_lib, times):
    from node_b200 import models, solver
    import torchdiffeq
    torch.manual_seed(4)
    f = models.ODEfunc(256).to(DEV)
    y0 = torch.randn(6, 256, 8, 8, device=DEV) * 0.5
    t = torch.tensor(times, device=DEV)
    with torch.no_grad():
        f.nfe = 0
        got = torchdiffeq.odeint(f, y0, t, rtol=1e-3, atol=1e-3, method='dopri5')
        st = dict(solver.last_stats)
        nfe_native = f.nfe
        os.environ['NODE_B200_WIDE'] = '0'
        try:
            f.nfe = 0
            ref = torchdiffeq.odeint(f, y0, t, rtol=1e-3, atol=1e-3, method='dopri5')
            st_ref = dict(solver.last_stats)
        finally:
            os.environ.pop('NODE_B200_WIDE', None)
    assert st['route'] == 'native-wide' and st_ref['route'] == 'generic'
    assert nfe_native == f.nfe == st['nfe'] and st['n_accept'] == st_ref['n_accept'] and st['n_reject'] == st_ref['n_reject']
    assert rel(got, ref) <= 1e-4, rel(got, ref)


def test_wide_odenet_forward(native_lib):
    from node_b200 import models, solver
    torch.manual_seed(0)
    net = models.ODENet(3, n_filters=256, downsample='residual', tol=1e-3).eval().to(DEV)
    x = torch.rand(8, 3, 32, 32, device=DEV)
    with torch.no_grad():
        got = net(x)
        route, nfe = solver.last_stats['route'], net.nfe(reset=True)
        os.environ['NODE_B200_WIDE'] = '0'
        try:
            ref = net(x)
        finally:
            os.environ.pop('NODE_B200_WIDE', None)
    assert route == 'native-wide' and nfe == net.nfe(reset=True)
    assert rel(got, ref) <= 1e-4 and bool((got.argmax(1) == ref.argmax(1)).all())


def _autograd_augmented(f, t, y, adj, tsign):
    """adjoint.py:32-55 on the eager module in float64; for a reversed span the SOLVER negates the augmented system and its time
    argument (misc.py:184-187): tsign * aug(tsign * t, .)."""
    import copy
    f64 = copy.deepcopy(f).double()
    params = tuple(f64.parameters())
    with torch.enable_grad():
        tt = torch.tensor(tsign * t, dtype=torch.float64, device=DEV, requires_grad=True)
        yy = y.double().requires_grad_(True)
        fe = f64(tt, yy)
        g = torch.autograd.grad(fe, (tt, yy) + params, -adj.double())
    return tsign * fe.detach(), tsign * g[1], tsign * g[0], tsign * torch.cat([q.reshape(-1) for q in g[2:]])


@pytest.mark.parametrize('C,hw,n,tsign', [(128, 8, 3, 1), (256, 8, 5, 1), (256, 8, 4, -1), (128, 7, 2, -1), (192, 8, 2, 1), (256, 8, 70, 1),
                                          (256, 8, 300, -1), (128, 8, 129, 1), (256, 7, 33, 1)])
def test_wide_augmented_dynamics(native_lib, C, hw, n, tsign):
    """One evaluation of the adjoint's augmented dynamics of a wide ODEfunc (node_b200.wide.WideAugmented: GroupNorm backward, data
    gradients and weight gradients on this repo's kernels) against float64 autograd of the eager module."""
    from node_b200 import models, wide
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
        aug.eval_into(torch.tensor(t, device=DEV), (y, adj), dst, float(tsign))
    ref = _autograd_augmented(f, t, y, adj, tsign)
    names = ('f', 'vjp_y', 'vjp_t', 'vjp_params')
    for name, got, want in zip(names, dst, ref):
        err = float((got.double() - want).abs().max() / want.abs().max())
        # A ReLU whose input is within rounding of zero opens in one precision and not in the other: with > 1e6 activations (the
        # large-batch case) one or two such flips are expected (tools/wide_vjp_debug.py: every stage up to the flipped GroupNorm is
        # at 1e-6, max error 6e-2 / L2 5e-4 after it). A flip reaches a 3x3 x C neighbourhood and its GroupNorm cells, so the check
        # there is the 99th percentile of the elementwise error instead of its maximum.
        if n >= 32 and got.numel() > 100:
            d = ((got.double() - want).abs() / want.abs().max()).reshape(-1)
            err = float(d.kthvalue(int(0.99 * d.numel())).values)
        # small batches: 5e-5 max-norm; large batch: vjp_t / vjp_params are sums over the batch (the flip is in every entry it touches)
        assert err <= (5e-5 if n < 32 else (5e-3 if got.numel() <= 100 else 5e-4)), (name, err)
    # per parameter tensor (a small tensor must not hide behind a large one)
    o = 0
    for pname, p in f.named_parameters():
        k = p.numel()
        got, want = dst[3][o:o + k].double(), ref[3][o:o + k]
        err = float((got - want).norm() / want.norm())
        assert err <= (1e-4 if n < 32 else 5e-3), (pname, err)
        o += k


@pytest.mark.parametrize('C,n', [(128, 4)])
def test_wide_adjoint_end_to_end(native_lib, monkeypatch, C, n):
    """odeint_adjoint of a wide ODEfunc: native augmented dynamics against the autograd / cuDNN route on the same solver kernels -
    identical NFE and step sequence, gradients within the conditioning of the reverse solve."""
    from node_b200 import models, odeint_adjoint, solver
    torch.manual_seed(C)
    f = models.ODEfunc(C).to(DEV)
    h0 = torch.randn(n, C, 8, 8, device=DEV)
    t = torch.tensor([0.0, 1.0], device=DEV)
    go = torch.randn(2, n, C, 8, 8, device=DEV)
    res = {}
    for mode in ('1', '0'):
        monkeypatch.setenv('NODE_B200_NATIVE_VJP', mode)
        h = h0.clone().requires_grad_(True)
        tt = t.clone().requires_grad_(True)
        for p in f.parameters():
            p.grad = None
        f.nfe = 0
        out = odeint_adjoint(f, h, tt, rtol=1e-3, atol=1e-3, method='dopri5')
        nfe_f, f.nfe = f.nfe, 0
        out.backward(go)
        st = dict(solver.last_stats)
        res[mode] = (h.grad.clone(), torch.cat([q.grad.reshape(-1) for q in f.parameters()]), tt.grad.clone(), nfe_f, f.nfe, st)
    a, b = res['1'], res['0']
    assert a[5]['adjoint_vjp'] == 'native-wide' and b[5]['adjoint_vjp'] == 'autograd'
    assert a[3] == b[3]
    same_seq = list(a[5]['trace']['accepted']) == list(b[5]['trace']['accepted'])
    # a 20+-step reverse solve has error ratios next to 1: a 1e-6 difference per evaluation may flip one accept / reject decision,
    # after which the two discretisations differ by the solver's own error (the golden test below pins the sequence on the reference)
    assert same_seq or abs(a[4] - b[4]) <= 12, (a[4], b[4])
    errs = [rel(a[i], b[i]) for i in range(3)]
    l2 = [float((a[i] - b[i]).norm() / b[i].norm()) for i in range(3)]
    print('wide adjoint C=%d: native vs autograd route  max-norm y0 %.2e params %.2e t %.2e | L2 y0 %.2e params %.2e | nfe %s' % (
        C, errs[0], errs[1], errs[2], l2[0], l2[1], a[3:5]))
    # two fp32 implementations of the same reverse solve: what separates them is its conditioning (tests/test_gpu_adjoint.py
    # measures 1e-2 .. 1e-1 max-norm on the reference's own arithmetic for 2-4 image batches), not the per-evaluation error (1e-6)
    assert max(errs) < 5e-2 and max(l2) < 1e-2, (errs, l2)


@pytest.mark.parametrize('name', ['adjoint_cifar_c128_n2', 'adjoint_cifar_c256_n2'])
def test_wide_adjoint_against_reference(native_lib, golden, name):
    """odeint_adjoint of a 128- / 256-filter ODE-Net block against the reference's own adjoint (tests/golden/adjoint_cifar_c*_n2.npz,
    tools/make_golden.py wide): forward, NFE forward / backward, backward accept / reject sequence (one rejected step) and dt trace,
    gradients within the conditioning the golden file records for the reference's own arithmetic."""
    from conftest import load_odefunc
    from node_b200 import odeint_adjoint, solver
    g = golden(name)
    func = load_odefunc(g, DEV).train()
    h0 = torch.from_numpy(g['h0']).to(DEV).requires_grad_(True)
    t = torch.from_numpy(g['t']).to(DEV).requires_grad_(True)
    tol = float(g['tol'])
    out = odeint_adjoint(func, h0, t, rtol=tol, atol=tol, method='dopri5')
    assert solver.last_stats['route'] == 'native-wide'
    nfe_f, func.nfe = func.nfe, 0
    out.backward(torch.from_numpy(g['grad_out']).to(DEV))
    st = dict(solver.last_stats)
    assert st.get('adjoint_vjp') == 'native-wide'
    assert nfe_f == int(g['nfe_f']) and func.nfe == int(g['nfe_b'])
    assert rel(out.detach().cpu(), torch.from_numpy(g['out'])) < 1e-4
    assert [bool(a) for a in st['trace']['accepted']] == [bool(a) for a in g['btr_acc']]
    dt_ref = torch.from_numpy(g['btr_dt'])
    assert float(((torch.tensor(list(st['trace']['dt']), dtype=torch.float64) - dt_ref).abs() / dt_ref.abs()).max()) < 1e-4
    gy, gt = h0.grad.cpu(), t.grad.cpu()
    gp = torch.cat([q.grad.reshape(-1) for q in func.parameters()]).cpu()
    cond_y = max(float(g['ref_err_y0']), float(g['sens_y0']))
    cond_p = max(float(g['ref_err_params']), float(g['sens_params']))
    ey, ep = rel(gy, torch.from_numpy(g['grad_y0'])), rel(gp, torch.from_numpy(g['grad_params']))
    et = float((gt - torch.from_numpy(g['grad_t'])).abs().max() / torch.from_numpy(g['grad_t']).abs().max())
    # one gate for all three: they come out of the same reverse trajectory, whose conditioning the golden file measured on the
    # reference's own arithmetic with two random 3e-6 perturbations of y(t1) (sens_*: the larger of the two is the scale)
    gate = max(1e-3, 2 * cond_y, 2 * cond_p)
    print('%s vs reference: y0 %.2e  params %.2e  t %.2e  (gate %.1e; ref sens y0 %.1e params %.1e)' % (name, ey, ep, et, gate, cond_y, cond_p))
    assert ey < gate and ep < gate and et < gate
