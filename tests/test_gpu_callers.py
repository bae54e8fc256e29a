"""GPU: the callers' GroupNorm -> ReLU pass (csrc/caller_ops.cu) against ATen's own group_norm + relu
(the reference's model.py:268-271 / nn.ReLU), and the whole ODENet forward with and without it."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('relu', [True, False])
@pytest.mark.parametrize('shape,groups', [((64, 64, 30, 30), 32), ((33, 64, 15, 15), 32), ((16, 64, 8, 8), 32), ((8, 24, 13, 13), 24),
                                          ((4, 64, 7, 7), 32), ((5, 64, 14, 14), 32), ((3, 256, 8, 8), 32), ((2, 64, 16, 16), 32)])
def test_groupnorm_relu_matches_aten(native_lib, shape, groups, relu):
    from node_b200 import caller_ops
    torch.manual_seed(1)
    norm = nn.GroupNorm(groups, shape[1]).to(DEV)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.uniform_(-0.5, 0.5)
        x = (torch.randn(shape, device=DEV) * 2.0 + 0.7)
        ref = norm(x)
        ref = torch.relu(ref) if relu else ref
        got = caller_ops.group_norm_relu(norm, x, relu=relu)
    torch.cuda.synchronize()
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_groupnorm_relu_keeps_autograd_path(native_lib):
    from node_b200 import caller_ops
    norm = nn.GroupNorm(32, 64).to(DEV)
    x = torch.randn(4, 64, 8, 8, device=DEV, requires_grad=True)
    y = caller_ops.group_norm_relu(norm, x)          # gradients needed: the modules' own PyTorch ops
    y.sum().backward()
    assert x.grad is not None and norm.weight.grad is not None


@pytest.mark.parametrize('in_ch,size', [(3, 32), (1, 28)])
@pytest.mark.parametrize('downsample', ['residual', 'convolution', 'minimal'])
def test_odenet_forward_same_with_fused_callers(native_lib, monkeypatch, downsample, in_ch, size):
    """Whole forward (CIFAR- and MNIST-shaped) with the callers' kernels against the same model with every caller
    kernel switched off (the reference's own PyTorch ops): identical NFE and top-1, logits within 1e-4."""
    from node_b200 import models, caller_ops
    torch.manual_seed(0)
    net = models.ODENet(in_ch, n_filters=64, downsample=downsample, tol=1e-3).eval().to(DEV)
    x = torch.rand(16, in_ch, size, size, device=DEV)
    torch.backends.cudnn.allow_tf32 = False
    with torch.no_grad():
        net.nfe(reset=True)
        c0 = caller_ops.launches
        fused = net(x)
        assert caller_ops.launches > c0
        nfe_f = net.nfe(reset=True)
        monkeypatch.setattr(caller_ops, '_fusable', lambda norm, x: False)
        monkeypatch.setattr(caller_ops, '_resconv_ok', lambda *a: False)
        monkeypatch.setattr(caller_ops, '_convs2_ok', lambda *a: False)
        monkeypatch.setattr(caller_ops, '_stem_ok', lambda *a: False)
        monkeypatch.setattr(models, 'head', lambda seq, x: seq(x))
        c1 = caller_ops.launches
        plain = net(x)
        assert caller_ops.launches == c1                     # nothing but PyTorch ops outside the ODE block
        nfe_p = net.nfe(reset=True)
    assert nfe_f == nfe_p
    assert float((fused - plain).abs().max()) <= 1e-4 * float(plain.abs().max())
    assert torch.equal(fused.argmax(1), plain.argmax(1))


@pytest.mark.parametrize('shape', [(37, 64, 15, 15), (16, 64, 8, 8), (9, 64, 13, 13), (11, 64, 7, 7), (450, 64, 15, 15), (1000, 64, 8, 8)])
def test_resblock_tail_matches_aten(native_lib, shape):
    """conv2(relu(norm2(x))) + shortcut (model.py:156-178) in one tcgen05 kernel vs cuDNN fp32 + ATen."""
    from node_b200 import caller_ops
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(2)
    norm = nn.GroupNorm(32, 64).to(DEV)
    conv = nn.Conv2d(64, 64, 3, 1, 1, bias=False).to(DEV)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.uniform_(-0.5, 0.5)
        x = torch.randn(shape, device=DEV) * 1.5 + 0.3
        sc = torch.randn(shape, device=DEV)
        ref = conv(torch.relu(norm(x))) + sc
        assert caller_ops._resconv_ok(norm, conv, x, sc)
        got = caller_ops.res_conv(norm, conv, x, sc)
        torch.cuda.synchronize()
        assert float((got - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
        nxt = nn.GroupNorm(32, 64).to(DEV)
        nxt.weight.uniform_(0.5, 1.5); nxt.bias.uniform_(-0.5, 0.5)
        ref_n = torch.relu(nxt(ref))                # the following block's norm1 -> ReLU fused into the same kernel
        got_n = caller_ops.res_conv(norm, conv, x, sc, next_norm=nxt)
        assert float((got_n - ref_n).abs().max()) <= 3e-5 * float(ref_n.abs().max())
        conv.weight.mul_(2.0)                       # in-place parameter update: the weight tiles are re-packed
        ref2 = conv(torch.relu(norm(x))) + sc
        got2 = caller_ops.res_conv(norm, conv, x, sc)
        assert float((got2 - ref2).abs().max()) <= 2e-5 * float(ref2.abs().max())


@pytest.mark.parametrize('shape', [(37, 64, 30, 30), (16, 64, 15, 15), (9, 64, 26, 26), (11, 64, 13, 13), (300, 64, 30, 30), (1000, 64, 15, 15)])
def test_resblock_head_matches_aten(native_lib, shape):
    """conv1 (3x3 stride 2) and the 1x1 stride-2 shortcut of a strided ResBlock (model.py:156-178) in one tcgen05 kernel."""
    from node_b200 import caller_ops
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(3)
    norm = nn.GroupNorm(32, 64).to(DEV)
    conv = nn.Conv2d(64, 64, 3, 2, 1, bias=False).to(DEV)
    down = nn.Conv2d(64, 64, 1, 2, bias=False).to(DEV)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.uniform_(-0.5, 0.5)
        a = torch.relu(norm(torch.randn(shape, device=DEV) * 1.5 + 0.3))
        ref_c, ref_s = conv(a), down(a)
        assert caller_ops._convs2_ok(norm, conv, down, a)
        c, s = caller_ops.res_head(norm, conv, down, a)
        torch.cuda.synchronize()
        assert c.shape == ref_c.shape and s.shape == ref_s.shape
        assert float((c - ref_c).abs().max()) <= 2e-5 * float(ref_c.abs().max())
        assert float((s - ref_s).abs().max()) <= 2e-5 * float(ref_s.abs().max())


@pytest.mark.parametrize('shape', [(37, 3, 32, 32), (19, 1, 28, 28), (300, 3, 32, 32)])
def test_stem_conv_groupnorm_relu_matches_aten(native_lib, shape):
    from node_b200 import caller_ops
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(4)
    conv = nn.Conv2d(shape[1], 64, 3, 1).to(DEV)
    norm = nn.GroupNorm(32, 64).to(DEV)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.uniform_(-0.5, 0.5)
        x = torch.rand(shape, device=DEV)
        ref = torch.relu(norm(conv(x)))
        assert caller_ops._stem_ok(conv, norm, x)
        got = caller_ops.stem_gn_relu(conv, norm, x)
        torch.cuda.synchronize()
        assert got.shape == ref.shape
        assert float((got - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


@pytest.mark.parametrize('features', [False, True])
@pytest.mark.parametrize('shape', [(37, 64, 8, 8), (5, 64, 7, 7), (3, 64, 16, 16), (2, 64, 6, 6)])
def test_classifier_head_matches_aten(native_lib, shape, features):
    """FCClassifier (model.py:231-250) - GroupNorm -> ReLU -> average pool -> Linear, and the feature-mode variant that
    stops after the pool (model.py:39-40, to_features_extractor) - in one pass."""
    from node_b200 import models, caller_ops
    torch.manual_seed(5)
    clf = models.FCClassifier(64, 10, dropout=0.5).eval().to(DEV)
    if features:
        clf.module[-1] = nn.Sequential()
    with torch.no_grad():
        clf.module[0].weight.uniform_(0.5, 1.5)
        clf.module[0].bias.uniform_(-0.5, 0.5)
        x = torch.randn(shape, device=DEV) * 1.5 + 0.3
        ref = clf.module(x)
        c0 = caller_ops.launches
        got = clf(x)
        torch.cuda.synchronize()
        assert caller_ops.launches == c0 + 1                 # the fused kernel ran
        assert got.shape == ref.shape
        assert float((got - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_weight_tile_caches_follow_the_live_parameters(native_lib, monkeypatch):
    """Models built, used and freed one after the other in one process (evaluate.py walks over runs): the packed weight
    tiles must always belong to the live parameters, also when addresses / ids are recycled."""
    import gc
    from node_b200 import models, caller_ops
    torch.backends.cudnn.allow_tf32 = False
    x = torch.rand(8, 3, 32, 32, device=DEV)
    outs = []
    for seed in (0, 1, 2, 1):
        torch.manual_seed(seed)
        net = models.ODENet(3, n_filters=64, downsample='residual', tol=1e-3).eval().to(DEV)
        with torch.no_grad():
            got = net(x)
            with monkeypatch.context() as m:
                m.setattr(caller_ops, '_fusable', lambda norm, x: False)
                m.setattr(caller_ops, '_resconv_ok', lambda *a: False)
                m.setattr(caller_ops, '_convs2_ok', lambda *a: False)
                m.setattr(caller_ops, '_stem_ok', lambda *a: False)
                plain_down = net.downsample(x)
            assert float((net.downsample(x) - plain_down).abs().max()) <= 1e-4 * float(plain_down.abs().max())
        outs.append(got.clone())
        del net
        gc.collect()
        torch.cuda.empty_cache()
    assert torch.equal(outs[1], outs[3])                     # same seed, same weights, same logits
    assert not torch.equal(outs[0], outs[1])
