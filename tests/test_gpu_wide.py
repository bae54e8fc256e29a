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


@pytest.mark.parametrize('C,n', [(256, 1), (256, 5), (256, 8), (128, 3), (256, 601), (128, 594)])
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
