"""GPU: backward of the callers (csrc/caller_bwd.cu + the raw-conv / weight-gradient engines) against ATen / cuDNN
fp32 autograd of the reference's own ops (model.py:119-178), piece by piece and through the whole training graph."""
import os

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(autouse=True)
def _fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize('relu', [True, False])
@pytest.mark.parametrize('shape', [(5, 64, 30, 30), (7, 64, 15, 15), (9, 64, 8, 8), (3, 64, 7, 7)])
def test_groupnorm_relu_backward_matches_aten(native_lib, shape, relu):
    from node_b200 import caller_grad
    torch.manual_seed(2)
    norm = nn.GroupNorm(32, 64).to(DEV)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.uniform_(-0.5, 0.5)
    x = (torch.randn(shape, device=DEV) * 2.0 + 0.3).requires_grad_(True)
    gy = torch.randn(shape, device=DEV)
    y = norm(x)
    y = torch.relu(y) if relu else y
    rx, rw, rb = torch.autograd.grad(y, [x, norm.weight, norm.bias], gy)
    gx, gw, gb = caller_grad.gn_relu_backward(x.detach(), gy, norm.weight.detach(), norm.bias.detach(), 32, norm.eps, relu)
    torch.cuda.synchronize()
    assert rel(gx, rx) <= 2e-5 and rel(gw, rw) <= 2e-5 and rel(gb, rb) <= 2e-5, (rel(gx, rx), rel(gw, rw), rel(gb, rb))


@pytest.mark.parametrize('hw', [30, 15])
def test_plane_split_merge(native_lib, hw):
    from node_b200 import caller_grad
    a = torch.randn(3, 64, hw, hw, device=DEV)
    planes = caller_grad.plane_split(a)
    ho = (hw - 1) // 2 + 1
    for pr in range(2):
        for pc in range(2):
            ref = torch.zeros(3, 64, ho, ho, device=DEV)
            sub = a[:, :, pr::2, pc::2]
            ref[:, :, :sub.shape[2], :sub.shape[3]] = sub
            assert torch.equal(planes[2 * pr + pc], ref)
    assert torch.equal(caller_grad.plane_merge(planes, hw, hw), a)


@pytest.mark.parametrize('n,hw', [(5, 15), (7, 8), (300, 8), (160, 15)])
def test_conv3x3_raw_matches_cudnn(native_lib, n, hw):
    from node_b200 import caller_grad
    torch.manual_seed(3)
    w = torch.randn(64, 64, 3, 3, device=DEV) * 0.05
    x = torch.randn(n, 64, hw, hw, device=DEV) * torch.rand(n, 1, 1, 1, device=DEV) * 1e-3        # gradient-like magnitudes
    add = torch.randn(n, 64, hw, hw, device=DEV) * 1e-4
    ref = F.conv2d(x.double(), w.double(), padding=1)
    got = caller_grad.conv3x3_raw(x, w)
    got2 = caller_grad.conv3x3_raw(x, w, addend=add, slot=1)
    torch.cuda.synchronize()
    assert rel(got.double(), ref) <= 2e-5, rel(got.double(), ref)
    assert rel(got2.double(), ref + add.double()) <= 2e-5


@pytest.mark.parametrize('n,hw', [(5, 15), (7, 8), (300, 8), (100, 15)])
def test_conv_wgrad_matches_autograd(native_lib, n, hw):
    from node_b200 import caller_grad
    torch.manual_seed(4)
    acts = [torch.relu(torch.randn(n, 64, hw, hw, device=DEV)) * 3.0 for _ in range(2)]
    grads = [torch.randn(n, 64, hw, hw, device=DEV) * 1e-3 for _ in range(2)]
    scale = torch.tensor([2.0 ** 10], device=DEV)          # |act| < 32: act * 2^10 < 2^15
    dw = caller_grad.conv_wgrad([acts[0], acts[1], acts[0]], [grads[0], grads[1], grads[1]], [scale, scale, scale])
    torch.cuda.synchronize()
    for p, (a, g) in enumerate([(acts[0], grads[0]), (acts[1], grads[1]), (acts[0], grads[1])]):
        w = torch.zeros(64, 64, 3, 3, device=DEV, dtype=torch.float64, requires_grad=True)
        ref, = torch.autograd.grad(F.conv2d(a.double(), w, padding=1), w, g.double())
        assert rel(dw[p].double(), ref) <= 2e-5, (p, rel(dw[p].double(), ref))


def _blocks(seed=0):
    from node_b200 import models
    torch.manual_seed(seed)
    net = models.ODENet(3, n_filters=64, downsample='residual', tol=1e-3, adjoint=True).to(DEV).train()
    return net


def _grads(net, x, y, native):
    os.environ['NODE_B200_CALLER_GRAD'] = '1' if native else '0'
    try:
        net.zero_grad(set_to_none=True)
        out = net.downsample(x)
        loss = (out * y).sum()
        loss.backward()
        return out.detach().clone(), {k: p.grad.detach().clone() for k, p in net.downsample.named_parameters()}
    finally:
        os.environ.pop('NODE_B200_CALLER_GRAD', None)


@pytest.mark.parametrize('n', [6, 150, 700])
def test_residual_downsampler_gradients_match_autograd(native_lib, n):
    """Every parameter gradient of the CIFAR residual downsampler, native forward + backward, against the modules' own ops in
    FLOAT64 (the truth) - ATen / cuDNN fp32 autograd is itself 1e-4..1e-3 away from it on the ill-conditioned batch sums
    (bias / beta / stem weight gradients), which the test records."""
    import copy
    from node_b200 import caller_grad
    net = _blocks()
    torch.manual_seed(5)
    x = torch.rand(n, 3, 32, 32, device=DEV)
    y = torch.randn(n, 64, 8, 8, device=DEV) * 1e-2
    before = caller_grad.launches
    out_n, g_n = _grads(net, x, y, True)
    assert caller_grad.launches > before, 'the native training path did not run'
    out_r, g_r = _grads(net, x, y, False)
    net64 = copy.deepcopy(net).double()
    out_t, g_t = _grads(net64, x.double(), y.double(), False)
    assert rel(out_n.double(), out_t) <= 1e-4
    worst = {k: rel(g_n[k].double(), g_t[k]) for k in g_t}
    aten = {k: rel(g_r[k].double(), g_t[k]) for k in g_t}
    print('n=%d native vs f64: max %.2e   ATen fp32 vs f64: max %.2e' % (n, max(worst.values()), max(aten.values())))
    # the composed graph is ill-conditioned in fp32 (ReLU masks flip on 1e-6 forward differences, the batch sums cancel): the
    # gate is "as close to float64 as ATen's own fp32 autograd, parameter by parameter"; the kernels themselves are gated at
    # 2e-5 in the per-piece tests above and in test_stem_backward_matches_float64
    bad = {k: (worst[k], aten[k]) for k in worst if worst[k] > max(1e-4, 3.0 * aten[k])}
    assert not bad, bad


@pytest.mark.parametrize('n', [5, 150, 700])
def test_stem_backward_matches_float64(native_lib, n):
    from node_b200 import caller_ops
    torch.manual_seed(8)
    conv0 = nn.Conv2d(3, 64, 3, 1).to(DEV)
    norm = nn.GroupNorm(32, 64).to(DEV)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.uniform_(-0.5, 0.5)
    x = torch.rand(n, 3, 32, 32, device=DEV)
    # positive upstream gradient: the batch sums do not cancel, so ONE ReLU mask that flips on a 1e-7 forward difference (it
    # happens from ~700 images on, to ATen's fp32 autograd as well: tools/stem_debug.py) cannot move a gradient by 1e-3
    go = (0.5 + torch.rand(n, 64, 30, 30, device=DEV)) * 1e-3
    ps = [conv0.weight, conv0.bias, norm.weight, norm.bias]
    out = caller_ops.stem_gn_relu(conv0, norm, x)
    assert type(out.grad_fn).__name__ == 'StemBackward'
    got = torch.autograd.grad(out, ps, go)
    c64, n64 = nn.Conv2d(3, 64, 3, 1).to(DEV).double(), nn.GroupNorm(32, 64).to(DEV).double()
    c64.load_state_dict({k: v.double() for k, v in conv0.state_dict().items()})
    n64.load_state_dict({k: v.double() for k, v in norm.state_dict().items()})
    ref = torch.autograd.grad(torch.relu(n64(c64(x.double()))), [c64.weight, c64.bias, n64.weight, n64.bias], go.double())
    aten = torch.autograd.grad(torch.relu(norm(conv0(x))), ps, go)
    errs = [rel(g.double(), r) for g, r in zip(got, ref)]
    errs_aten = [rel(g.double(), r) for g, r in zip(aten, ref)]
    # fp32 forward differences of 1e-7 flip a handful of ReLU masks in 40M elements and the batch sums cancel: beyond 2e-5 the
    # gate is ATen's own distance from float64
    assert all(e <= max(2e-5, 3.0 * a) for e, a in zip(errs, errs_aten)), (errs, errs_aten)


def test_mnist_residual_downsampler_gradients(native_lib):
    """MNIST shapes (28 -> 26 -> 13 -> 7): the same native backward through the 13x13 / 7x7 engines, against float64."""
    import copy
    from node_b200 import caller_grad, models
    torch.manual_seed(3)
    net = models.ODENet(1, n_filters=64, downsample='residual', tol=1e-3, adjoint=True).to(DEV).train()
    x = torch.rand(40, 1, 28, 28, device=DEV)
    y = (0.5 + torch.rand(40, 64, 7, 7, device=DEV)) * 1e-2
    before = caller_grad.launches
    out_n, g_n = _grads(net, x, y, True)
    assert caller_grad.launches - before > 40, 'the native training path did not run on every block'
    out_r, g_r = _grads(net, x, y, False)
    out_t, g_t = _grads(copy.deepcopy(net).double(), x.double(), y.double(), False)
    worst = {k: rel(g_n[k].double(), g_t[k]) for k in g_t}
    aten = {k: rel(g_r[k].double(), g_t[k]) for k in g_t}
    bad = {k: (worst[k], aten[k]) for k in worst if worst[k] > max(1e-4, 3.0 * aten[k])}
    assert not bad, bad


def test_training_step_gradients_whole_model(native_lib):
    """cfg3: forward + CE loss + adjoint backward; all parameter gradients with the native caller backward vs PyTorch's."""
    net = _blocks(1)
    torch.manual_seed(6)
    x = torch.rand(32, 3, 32, 32, device=DEV)
    y = torch.randint(0, 10, (32,), device=DEV)

    def run(native):
        os.environ['NODE_B200_CALLER_GRAD'] = '1' if native else '0'
        try:
            net.zero_grad(set_to_none=True)
            loss = F.cross_entropy(net(x), y)
            loss.backward()
            return float(loss), {k: p.grad.detach().clone() for k, p in net.named_parameters()}
        finally:
            os.environ.pop('NODE_B200_CALLER_GRAD', None)
    net.classifier.module[3].p = 0.0 if isinstance(net.classifier.module[3], nn.Dropout) else None
    l_n, g_n = run(True)
    l_r, g_r = run(False)
    assert abs(l_n - l_r) <= 1e-5 * abs(l_r)
    worst = {k: rel(g_n[k], g_r[k]) for k in g_r}
    # both arms share the adjoint ODE block; the callers' fp32 ATen gradients are themselves ~1e-3 from float64 (see above)
    assert max(worst.values()) <= 3e-2, worst
