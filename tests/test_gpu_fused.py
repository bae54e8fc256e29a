"""GPU: the fused sm_100a route (recognised ODEfunc) against the reference's golden vectors -
identical NFE and accept/reject sequence, dt trace to 1e-5, outputs to 1e-4 relative (BASELINE
north_star), top-1 identical - through the reference-facing API (ODENet / ODEBlock / odeint)."""
import numpy as np
import pytest
import torch

from conftest import load_odefunc, odefunc_params
from oracle import dopri5_port, odefunc_port

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL_OUT = 1e-4          # fp32 contract of BASELINE.json north_star
CASES = ['cifar_res_n8', 'cifar_res_n7_tol1e-4', 'mnist_conv_n9', 'mnist_res_n5', 'cifar_oneshot_n3', 'mnist_oneshot_n3']


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize('mode', ['simt', 'tf32x3', 'f16x3'])
@pytest.mark.parametrize('name', CASES)
def test_fused_dynamics_kernel(native_lib, golden, monkeypatch, name, mode):
    from node_b200 import solver
    monkeypatch.setenv('NODE_B200_CONV', mode)
    g = golden(name)
    func = load_odefunc(g, DEV)
    h0 = torch.from_numpy(g['h0']).to(DEV)
    k = solver.odefunc_forward(func, 0.37, h0)
    torch.cuda.synchronize()
    assert rel(k.cpu(), torch.from_numpy(g['f037'])) < 2e-5
    # reversed-time wrapper: -f(-t, y)
    p = odefunc_params(g)
    kr = solver.odefunc_forward(func, 0.25, h0, tsign=-1.0)
    ref = -odefunc_port.odefunc_forward(p, torch.tensor(-0.25), torch.from_numpy(g['h0']))
    assert rel(kr.cpu(), ref) < 2e-5


@pytest.mark.parametrize('name', ['cifar_res_n8', 'mnist_conv_n9', 'cifar_oneshot_n3'])
def test_fused_dynamics_ill_conditioned_groupnorm(native_lib, golden, name):
    """States whose GroupNorm cells have |mean| >> std: the step engine's one-pass moments must fall back to the
    two-pass variance of native_group_norm (model.py:268-271) instead of cancelling."""
    from node_b200 import solver
    g = golden(name)
    func = load_odefunc(g, DEV)
    p = odefunc_params(g)
    h0 = torch.from_numpy(g['h0'])
    gen = torch.Generator().manual_seed(3)
    for offset, noise, tol in [(30.0, 1.0, 1e-4), (1000.0, 0.5, 2e-3), (5.0, 0.0, 5e-4)]:   # fp32 itself is ~2e-4 / 1e-4 from fp64 on the last two
        y = offset * torch.sign(torch.randn(h0.shape[0], h0.shape[1], 1, 1, generator=gen)) + noise * h0
        y[0] = h0[0]                                   # well- and ill-conditioned images share a super-tile
        ref = odefunc_port.odefunc_forward(p, torch.tensor(0.37), y)
        k = solver.odefunc_forward(func, 0.37, y.to(DEV))
        torch.cuda.synchronize()
        assert torch.isfinite(k).all()
        assert rel(k.cpu(), ref) < tol, (offset, noise)
        assert rel(k[0].cpu(), ref[0]) < 2e-5


def test_tf32_single_pass_mode_is_reported_separately(native_lib, golden, monkeypatch):
    from node_b200 import solver
    monkeypatch.setenv('NODE_B200_CONV', 'tf32')
    g = golden('cifar_res_n8')
    k = solver.odefunc_forward(load_odefunc(g, DEV), 0.37, torch.from_numpy(g['h0']).to(DEV))
    e = rel(k.cpu(), torch.from_numpy(g['f037']))
    assert 2e-5 < e < 5e-3, e        # 1xTF32 is outside the fp32 contract, by about this much


@pytest.mark.parametrize('mode', ['simt', 'tf32x3', 'f16x3'])
@pytest.mark.parametrize('name', CASES)
def test_fused_solve_matches_reference(native_lib, golden, monkeypatch, name, mode):
    from node_b200 import odeint, solver
    monkeypatch.setenv('NODE_B200_CONV', mode)
    g = golden(name)
    func = load_odefunc(g, DEV)
    h0, t, tol = torch.from_numpy(g['h0']).to(DEV), torch.from_numpy(g['t']).to(DEV), float(g['tol'])
    with torch.no_grad():
        out = odeint(func, h0, t, rtol=tol, atol=tol, method='dopri5')
    st = dict(solver.last_stats)
    assert st['route'] == 'fused' and st['status'] == 0
    assert st['nfe'] == int(g['nfe']) == func.nfe                   # identical NFE, counted on the module
    assert list(st['trace']['accepted']) == list(g['tr_acc'])       # identical accept/reject sequence
    np.testing.assert_allclose(st['trace']['dt'], g['tr_dt'], rtol=1e-5)
    np.testing.assert_allclose(st['trace']['t'], g['tr_t'], rtol=1e-5, atol=1e-12)
    assert rel(out.cpu(), torch.from_numpy(g['out'])) < TOL_OUT
    assert torch.equal(out[0], h0)                                  # solvers.py:27


def test_fused_multi_time_dense_output(native_lib, golden):
    """ICMR feature extraction: 10 output times cost no extra evaluations (dopri5.py:85-92)."""
    from node_b200 import odeint, solver
    g = golden('cifar_res_n8_t10')
    func = load_odefunc(g, DEV)
    h0, t = torch.from_numpy(g['h0']).to(DEV), torch.from_numpy(g['t']).to(DEV)
    with torch.no_grad():
        out = odeint(func, h0, t, rtol=1e-3, atol=1e-3, method='dopri5')
    assert out.shape == (10,) + tuple(h0.shape)
    assert solver.last_stats['nfe'] == int(g['nfe']) == 26
    assert rel(out.cpu(), torch.from_numpy(g['out'])) < TOL_OUT
    for i in range(10):
        assert rel(out[i].cpu(), torch.from_numpy(g['out'][i])) < 2 * TOL_OUT


@pytest.mark.parametrize('name,in_ch,size,ds', [('cifar_res_n128', 3, 32, 'residual'), ('mnist_conv_n128', 1, 28, 'convolution')])
def test_odenet_end_to_end_batch128(native_lib, golden, name, in_ch, size, ds):
    """BASELINE configs 1/2 at the reference batch size, from seeds: logits, top-1, NFE, dt trace."""
    from node_b200 import models, solver
    g = golden(name)
    torch.manual_seed(int(g['seed']))
    net = models.ODENet(in_ch, n_filters=64, downsample=ds, tol=1e-3).eval()
    x = torch.rand(int(g['N']), in_ch, size, size)
    net, x = net.to(DEV), x.to(DEV)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False                        # fp32 contract for the downsampler convs
    try:
        with torch.no_grad():
            logits = net(x)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    st = dict(solver.last_stats)
    ref = torch.from_numpy(g['logits'])
    assert st['route'] == 'fused' and net.nfe() == int(g['nfe'])
    assert list(st['trace']['accepted']) == list(g['tr_acc'])
    np.testing.assert_allclose(st['trace']['dt'], g['tr_dt'], rtol=1e-5)
    assert rel(logits.cpu(), ref) < TOL_OUT
    assert torch.equal(logits.argmax(1).cpu(), ref.argmax(1))       # top-1 identical


def test_fused_reversed_time(native_lib, golden):
    """Decreasing t integrates -f(-t, y) (misc.py:184-187); checked against the oracle."""
    from node_b200 import odeint, solver
    g = golden('cifar_res_n8')
    func = load_odefunc(g, DEV)
    p = odefunc_params(g)
    y1 = torch.from_numpy(g['out'][-1])
    t = torch.tensor([1.0, 0.5])
    tr = dopri5_port.Trace()
    with torch.no_grad():
        ref = dopri5_port.dopri5_solve(lambda a, b: odefunc_port.odefunc_forward(p, a, b), y1, t, 1e-3, 1e-3, trace=tr)
        out = odeint(func, y1.to(DEV), t.to(DEV), rtol=1e-3, atol=1e-3, method='dopri5')
    st = solver.last_stats
    assert st['route'] == 'fused' and st['nfe'] == tr.nfe
    assert list(st['trace']['accepted']) == [s[2] for s in tr.steps]
    assert rel(out.cpu(), ref) < TOL_OUT


# odeint_adjoint end to end: tests/test_gpu_adjoint.py (every shape, gates derived from the reference's own fp32-vs-fp64 distance)


def test_training_without_adjoint_flag(native_lib, monkeypatch):
    """train.py's default (no --adjoint, model.py:359) asks odeint itself for gradients: served by the unrolled route (the
    reference's own gradient, tests/test_gpu_unrolled.py); NODE_B200_ODEINT_GRAD=adjoint serves it by the adjoint ODE, which then
    equals the --adjoint model's gradient. The whole ODENet trains either way (finite gradients for every parameter)."""
    import warnings
    from node_b200 import models, solver
    grads = {}
    for mode, adjoint in (('unrolled', False), ('adjoint', False), ('adjoint', True)):
        monkeypatch.setenv('NODE_B200_ODEINT_GRAD', mode)
        torch.manual_seed(0)
        net = models.ODENet(3, n_filters=64, downsample='residual', tol=1e-3, adjoint=adjoint).train().to(DEV)
        x = torch.rand(8, 3, 32, 32, generator=torch.Generator().manual_seed(5)).to(DEV)
        y = torch.randint(0, 10, (8,), generator=torch.Generator().manual_seed(6)).to(DEV)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            loss = torch.nn.functional.cross_entropy(net(x), y)
        loss.backward()
        if mode == 'unrolled':                 # recognised dynamics: fused forward, the unrolled graph recorded inside backward
            assert solver.last_stats['route'] == 'fused' and solver.last_stats.get('grad_route') == 'unrolled'
        grads[(mode, adjoint)] = [p.grad.clone() for p in net.parameters()]
        assert all(g is not None and torch.isfinite(g).all() for g in grads[(mode, adjoint)])
    for a, b in zip(grads[('adjoint', False)], grads[('adjoint', True)]):      # same kernels; cuDNN's backward is not bit-reproducible
        assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max()) + 1e-9
    # unrolled vs adjoint: two DIFFERENT gradients of the same loss at tol 1e-3 (the reference's own two are 1.2e-1 .. 5.5e-1 apart in
    # max norm on the golden problems, tests/golden/unrolled_*.npz `adjoint_dev_*`) that point the same way (cosine 0.98 here)
    va = torch.cat([g.reshape(-1) for g in grads[('unrolled', False)]])
    vb = torch.cat([g.reshape(-1) for g in grads[('adjoint', True)]])
    assert float(torch.dot(va, vb) / (va.norm() * vb.norm())) > 0.95


def test_cuda_graph_replay_of_the_solve(native_lib, golden, monkeypatch):
    """north_star (3): the launch sequence of a solve is captured in a CUDA graph and replayed (launch-bound sizes)."""
    from node_b200 import odeint, solver
    g = golden('cifar_res_n8')
    func = load_odefunc(g, DEV)
    h0, t, tol = torch.from_numpy(g['h0']).to(DEV), torch.from_numpy(g['t']).to(DEV), float(g['tol'])
    h1 = h0.flip(0).contiguous() * 0.9
    outs = {}
    for mode in ('0', '1'):
        monkeypatch.setenv('NODE_B200_GRAPH', mode)
        solver._step_guess.clear(); solver._graph_cache.clear()
        with torch.no_grad():
            a = odeint(func, h0, t, rtol=tol, atol=tol, method='dopri5')      # first call: direct (learns the step count)
            b = odeint(func, h0, t, rtol=tol, atol=tol, method='dopri5')      # second call: replayed when mode == '1'
            st = dict(solver.last_stats)
            c = odeint(func, h1, t, rtol=tol, atol=tol, method='dopri5')      # same graph, new input
        assert (len(solver._graph_cache) > 0) == (mode == '1')
        assert torch.equal(a, b) and st['nfe'] == int(g['nfe']) and st['status'] == 0
        assert list(st['trace']['accepted']) == list(g['tr_acc'])
        assert rel(b.cpu(), torch.from_numpy(g['out'])) < TOL_OUT
        assert torch.equal(c[0], h1)
        outs[mode] = (b.clone(), c.clone())
    assert torch.equal(outs['0'][0], outs['1'][0]) and torch.equal(outs['0'][1], outs['1'][1])


@pytest.mark.parametrize('tol', [1e-4, 1e-2, 1e-1])
def test_tolerance_sweep_matches_oracle(native_lib, golden, tol):
    """cfg5's sweep (adversarial/reproduce.sh:9-10: tol 1e-4 ... 1e-1) on a small batch: NFE, accept/reject sequence and
    outputs against the oracle evaluated here on the CPU."""
    from node_b200 import odeint, solver
    g = golden('cifar_res_n8')
    func = load_odefunc(g, DEV)
    p = odefunc_params(g)
    h0 = torch.from_numpy(g['h0'])[:4].contiguous()
    t = torch.tensor([0., 1.])
    tr = dopri5_port.Trace()
    ref = dopri5_port.dopri5_solve(lambda tt, y: odefunc_port.odefunc_forward(p, tt, y), h0, t, tol, tol, trace=tr)
    with torch.no_grad():
        out = odeint(func, h0.to(DEV), t.to(DEV), rtol=tol, atol=tol, method='dopri5')
    st = dict(solver.last_stats)
    assert st['route'] == 'fused' and st['nfe'] == tr.nfe
    assert [int(a) for a in st['trace']['accepted']] == [int(s[2]) for s in tr.steps]
    assert rel(out.cpu(), ref) < TOL_OUT


@pytest.mark.parametrize('n', [445, 900, 1337])
def test_two_slot_step_kernel_matches_oracle(native_lib, golden, n):
    """Batches above 444 images run the two-slot variant of the step kernel (NST > 148 super-tiles), including its tail
    rounds where only slot 0 has work: one evaluation and a whole solve against the oracle on the CPU."""
    from node_b200 import odeint, solver
    g = golden('cifar_res_n8')
    func = load_odefunc(g, DEV)
    p = odefunc_params(g)
    base = torch.from_numpy(g['h0'])
    gen = torch.Generator().manual_seed(n)
    reps = (n + base.shape[0] - 1) // base.shape[0]
    h0 = (base.repeat(reps, 1, 1, 1)[:n] * (0.8 + 0.4 * torch.rand(n, 1, 1, 1, generator=gen))).contiguous()
    ref_k = odefunc_port.odefunc_forward(p, torch.tensor(0.37), h0)
    k = solver.odefunc_forward(func, 0.37, h0.to(DEV))
    assert rel(k.cpu(), ref_k) < 2e-5
    t = torch.tensor([0., 1.])
    tr = dopri5_port.Trace()
    ref = dopri5_port.dopri5_solve(lambda tt, y: odefunc_port.odefunc_forward(p, tt, y), h0, t, 1e-3, 1e-3, trace=tr)
    with torch.no_grad():
        out = odeint(func, h0.to(DEV), t.to(DEV), rtol=1e-3, atol=1e-3, method='dopri5')
    st = dict(solver.last_stats)
    assert st['route'] == 'fused' and st['nfe'] == tr.nfe
    assert [int(a) for a in st['trace']['accepted']] == [int(s[2]) for s in tr.steps]
    assert rel(out.cpu(), ref) < TOL_OUT
    # per image, not only in the maximum norm
    per = (out[-1].cpu() - ref[-1]).flatten(1).abs().amax(1) / ref[-1].flatten(1).abs().amax(1)
    assert float(per.max()) < 2 * TOL_OUT


def test_cuda_graphs_survive_workspace_recycling(native_lib, golden, monkeypatch):
    """More shapes than the workspace cache holds: graphs captured against released workspaces must never be replayed
    (device addresses get recycled across shapes)."""
    from node_b200 import odeint, solver
    monkeypatch.setenv('NODE_B200_GRAPH', '1')
    g = golden('cifar_res_n8')
    func = load_odefunc(g, DEV)
    base = torch.from_numpy(g['h0']).to(DEV)
    t = torch.from_numpy(g['t']).to(DEV)
    tol = float(g['tol'])
    ref = {}
    with torch.no_grad():
        for rnd in range(2):
            for n in list(range(1, 9)) + [8, 7, 3, 1]:
                h0 = base[:n].contiguous()
                for _ in range(2):                                   # direct, then replayed
                    out = odeint(func, h0, t, rtol=tol, atol=tol, method='dopri5')
                assert out.shape == (len(t), n) + tuple(base.shape[1:])
                assert solver.last_stats['status'] == 0
                if n in ref:
                    assert torch.equal(out, ref[n])
                else:
                    ref[n] = out.clone()
            for shape in [(2, 64, 6, 6), (2, 64, 7, 7), (2, 64, 14, 14), (3, 64, 8, 8), (5, 64, 6, 6), (4, 64, 7, 7)]:   # evict
                y = torch.randn(shape, device=DEV)
                odeint(func, y, t, rtol=tol, atol=tol, method='dopri5')
                odeint(func, y, t, rtol=tol, atol=tol, method='dopri5')
