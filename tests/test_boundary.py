"""CPU: the drop-in boundary - C ABI exports, API surface and error behaviour of odeint /
odeint_adjoint (reference odeint.py:20-76, adjoint.py:105-133, misc.py:173-195), module mirror."""
import os
import re

import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, 'include', 'node_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(node_b200_\w+)\s*\(', src)))


def test_library_exports_every_declared_symbol(native_lib):
    from node_b200 import native
    names = declared_functions()
    assert len(names) >= 16
    for n in names:
        assert hasattr(native_lib, n), n
    assert sorted(native.EXPORTS) == names
    assert native_lib.node_b200_abi_version() == native.ABI_VERSION


def test_ctl_layout_is_consistent(native_lib):
    from node_b200 import native
    L = native.layout()
    assert L['max_seg'] == 8 and L['sizeof'] % 8 == 0
    assert L['t0'] == 0 and L['t1'] == 8 and L['dt'] == 16
    offs = [v for k, v in L.items() if k not in ('sizeof', 'partial_blocks', 'max_seg', 'max_trace')]
    assert max(offs) < L['sizeof'] and len(set(offs)) == len(offs)


def test_workspace_size_and_shape_support(native_lib):
    f = native_lib.node_b200_fused_workspace_bytes
    assert f(128, 64, 8, 8) > 10 * 128 * 64 * 64 * 4
    for hw in ((6, 6), (7, 7), (14, 14), (16, 16)):
        assert f(5, 64, *hw) > 0
    assert f(4, 64, 32, 32) < 0      # too large for a CTA-resident image: generic route
    assert f(4, 256, 8, 8) < 0       # 256 filters: generic route (out of scope this round)


def test_torchdiffeq_dropin_exports():
    import torchdiffeq
    import node_b200
    assert torchdiffeq.odeint is node_b200.odeint and torchdiffeq.odeint_adjoint is node_b200.odeint_adjoint
    import inspect
    assert str(inspect.signature(torchdiffeq.odeint)) == '(func, y0, t, rtol=1e-07, atol=1e-09, method=None, options=None)'
    assert str(inspect.signature(torchdiffeq.odeint_adjoint)) == '(func, y0, t, rtol=1e-06, atol=1e-12, method=None, options=None)'


def test_error_behaviour_matches_reference():
    from node_b200 import odeint, odeint_adjoint
    f = lambda t, y: -y
    y0, t = torch.ones(3), torch.tensor([0., 1.])
    with pytest.raises(ValueError):
        odeint(f, y0, t, options={'safety': 0.8})                 # odeint.py:65-66
    with pytest.raises(KeyError):
        odeint(f, y0, t, method='no-such-method')                 # odeint.py:71
    with pytest.raises(NotImplementedError):
        odeint(f, y0, t, method='rk4')                            # out of scope: only dopri5
    with pytest.raises(TypeError):
        odeint(f, torch.ones(3, dtype=torch.int64), t)            # misc.py:190-191
    with pytest.raises(TypeError):
        odeint(f, y0, torch.tensor([0, 1]))                       # misc.py:192-193
    with pytest.raises(ValueError):
        odeint_adjoint(f, y0, t)                                  # adjoint.py:109-110
    with pytest.raises(RuntimeError, match='CUDA-only'):
        odeint(f, y0, t)                                          # no CPU fallback, loudly
    with pytest.raises(AssertionError):
        odeint(f, [y0], t)                                        # misc.py:181


def test_no_silent_autograd_through_odeint(monkeypatch):
    """Gradients through `odeint` with a plain callable (api_tests.py:30-38, train.py without --adjoint) are never silently
    dropped: the default serves them by the unrolled route (node_b200.unrolled), NODE_B200_ODEINT_GRAD=adjoint by the adjoint ODE,
    where the modules the callable closes over are found so that their parameters are differentiated too. (No GPU here: either
    request must get as far as the CUDA-only check, not return a graph-less tensor or run on the CPU.)"""
    from node_b200 import odeint, solver
    lin = nn.Linear(3, 3)
    f = lambda t, y: lin(y)
    assert solver._closure_modules(f) == [lin]
    assert solver._needs_grad(f, (torch.ones(3),), torch.tensor([0., 1.]))        # lin's parameters require grad
    with pytest.raises(RuntimeError, match='CUDA-only'):
        odeint(f, torch.ones(3, requires_grad=True), torch.tensor([0., 1.]))
    monkeypatch.setenv('NODE_B200_ODEINT_GRAD', 'adjoint')
    with pytest.warns(UserWarning, match='adjoint'), pytest.raises(RuntimeError, match='CUDA-only'):
        odeint(f, torch.ones(3, requires_grad=True), torch.tensor([0., 1.]))


def test_recogniser_structure_checks():
    from node_b200 import models, solver as api
    f = models.ODEfunc(64)
    assert api.recognise_odefunc(f) is None                       # parameters not on a CUDA device
    assert api.recognise_odefunc(models.ODEfunc(64, norm='batch')) is None
    assert api.recognise_odefunc(nn.Linear(2, 2)) is None
    assert api.recognise_odefunc(lambda t, y: y) is None


def test_odeblock_mirror_semantics():
    from node_b200 import models
    blk = models.ODEBlock(64, t1=0)
    x = torch.randn(2, 64, 8, 8)
    assert blk(x) is x                                            # model.py:363-364
    blk.t1 = [0.25, 0.5, 1]
    assert blk.integration_time.tolist() == [0, 0.25, 0.5, 1]     # model.py:397-399 prepends 0
    blk.t1 = 2
    assert blk.integration_time.tolist() == [0, 2] and blk.integration_time.dtype == torch.float32
    with pytest.raises(ValueError):
        blk.t1 = 'x'
    net = models.ODENet(3, adjoint=True)
    from node_b200 import odeint_adjoint
    assert net.odeblock.odeint is odeint_adjoint
    keys = list(net.state_dict().keys())
    for k in ('odeblock.odefunc.norm1.weight', 'odeblock.odefunc.conv1._layer.weight', 'odeblock.odefunc.conv2._layer.bias',
              'odeblock.odefunc.norm3.bias', 'classifier.module.4.weight'):
        assert k in keys
    assert tuple(net.state_dict()['odeblock.odefunc.conv1._layer.weight'].shape) == (64, 65, 3, 3)
    assert sum(p.numel() for p in net.odeblock.odefunc.parameters()) == 75392


def test_missing_library_fails_loudly(monkeypatch):
    from node_b200 import native
    monkeypatch.setattr(native, '_lib', None)
    monkeypatch.setattr(native, 'LIB_PATH', '/nonexistent/libnode_b200.so')
    with pytest.raises(RuntimeError, match='no CPU or PyTorch fallback'):
        native.lib()


def test_caller_modules_keep_the_reference_ops_off_the_gpu_or_with_gradients():
    """node_b200.models routes the callers' hot spots to CUDA kernels only for CUDA tensors without autograd; on the CPU
    (and for training) ResDownsample / FCClassifier are the reference's PyTorch ops (model.py:119-178, 231-250)."""
    from node_b200 import models
    torch.manual_seed(0)
    for name in ('residual', 'convolution', 'minimal', 'one-shot'):
        net = models.ODENet(3, n_filters=64, downsample=name, tol=1e-3).eval()
        x = torch.rand(2, 3, 32, 32)
        with torch.no_grad():
            got = net.downsample(x)
            ref = net.downsample.module(x)                 # plain nn.Sequential forward
            assert torch.equal(got, ref)
            logits = net.classifier(got)
            assert torch.equal(logits, net.classifier.module(got))
    net = models.ODENet(3, n_filters=64, downsample='residual', tol=1e-3).train()
    x = torch.rand(2, 3, 32, 32, requires_grad=True)
    h = net.downsample(x)
    net.classifier(h).sum().backward()
    assert x.grad is not None and net.downsample.module[1].conv1.weight.grad is not None


def test_time_tensor_cache_and_invalidate():
    """Host copies of time tensors are cached by (identity, _version); a tensor that requires grad is never cached (gradcheck
    perturbs through `.data`, gradient_tests.py:33-37), and invalidate_caches() forgets the rest."""
    from node_b200 import invalidate_caches, solver
    t = torch.tensor([0., 1.])
    assert solver._host_times(t)[1] == 1.0
    t.data[1] = 2.0                                   # no version bump: the cached copy is stale by design ...
    assert solver._host_times(t)[1] == 1.0
    invalidate_caches()                               # ... until the caller says so
    assert solver._host_times(t)[1] == 2.0
    tg = torch.tensor([0., 1.], requires_grad=True)
    assert solver._host_times(tg)[1] == 1.0
    tg.data[1] = 3.0
    assert solver._host_times(tg)[1] == 3.0
