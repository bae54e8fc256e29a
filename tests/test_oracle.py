"""CPU: the oracle (torch restatement of the reference) against the committed golden vectors that
tools/make_golden.py recorded from the unmodified reference."""
import numpy as np
import pytest
import torch

from conftest import odefunc_params
from oracle import dopri5_port, odefunc_port

ODENET_CASES = ['cifar_res_n8', 'cifar_res_n8_t10', 'cifar_res_n7_tol1e-4', 'mnist_conv_n9', 'mnist_res_n5',
                'cifar_oneshot_n3', 'mnist_oneshot_n3']


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize('name', ODENET_CASES)
def test_odenet_solve_matches_reference(golden, name):
    g = golden(name)
    p = odefunc_params(g)
    h0, t, tol = torch.from_numpy(g['h0']), torch.from_numpy(g['t']), float(g['tol'])
    tr = dopri5_port.Trace()
    with torch.no_grad():
        out = dopri5_port.dopri5_solve(lambda a, b: odefunc_port.odefunc_forward(p, a, b), h0, t, tol, tol, trace=tr)
        f = odefunc_port.odefunc_forward(p, torch.tensor(0.37), h0)
    # conv kernels may differ between host CPUs, so allow fp32 noise on values but not on decisions
    assert rel(out, torch.from_numpy(g['out'])) < 2e-5
    assert rel(f, torch.from_numpy(g['f037'])) < 2e-5
    assert tr.nfe == int(g['nfe']) == 2 + 6 * len(tr.steps)
    assert [s[2] for s in tr.steps] == list(g['tr_acc'])
    np.testing.assert_allclose([s[1] for s in tr.steps], g['tr_dt'], rtol=1e-5)
    assert bool((out[0] == h0).all())


def test_time_map_fold_equals_concat_conv(golden):
    g = golden('cifar_res_n8')
    p = odefunc_params(g)
    h0 = torch.from_numpy(g['h0'])
    with torch.no_grad():
        a = odefunc_port.odefunc_forward(p, torch.tensor(0.8), h0)
        b = odefunc_port.odefunc_forward_folded(p, torch.tensor(0.8), h0)
    assert rel(b, a) < 5e-6
    tm = odefunc_port.time_map(p['conv1._layer.weight'], 8, 8)
    assert len(torch.unique(tm[0])) <= 9          # corner / edge / interior classes


def test_hand_derived_vjp_matches_autograd(golden):
    g = golden('adjoint_cifar_n4')
    p = {k: v.clone().requires_grad_(True) for k, v in odefunc_params(g).items()}
    x = torch.from_numpy(g['h0']).clone().requires_grad_(True)
    t = torch.tensor(0.4, requires_grad=True)
    a = torch.from_numpy(g['grad_out'][-1]) * 100
    f = odefunc_port.odefunc_forward(p, t, x)
    grads = torch.autograd.grad(f, [x, t] + [p[k] for k in odefunc_port.PARAM_ORDER], a)
    with torch.no_grad():
        f2, vx, vt, vp = odefunc_port.odefunc_vjp({k: v.detach() for k, v in p.items()}, t.detach(), x.detach(), a)
    assert rel(f2, f.detach()) < 1e-5
    assert rel(vx, grads[0]) < 1e-4
    assert abs(float(vt) - float(grads[1])) < 1e-4 * (abs(float(grads[1])) + 1e-3)
    assert rel(vp, torch.cat([q.reshape(-1) for q in grads[2:]])) < 1e-4


def test_adjoint_matches_reference(golden):
    g = golden('adjoint_cifar_n4')
    p = odefunc_params(g)
    params = [p[k] for k in odefunc_port.PARAM_ORDER]
    t, tol = torch.from_numpy(g['t']), float(g['tol'])
    tr = dopri5_port.Trace()
    gy, gt, gp = dopri5_port.adjoint_backward(
        lambda a, b: odefunc_port.odefunc_forward(p, a, b), params, t, torch.from_numpy(g['out']),
        torch.from_numpy(g['grad_out']), tol, tol, trace=tr,
        vjp=lambda a, b, c: odefunc_port.odefunc_vjp(p, a, b, c))
    assert rel(gy, torch.from_numpy(g['grad_y0'])) < 1e-4
    assert rel(gp, torch.from_numpy(g['grad_params'])) < 1e-4
    assert [s[2] for s in tr.steps] == list(g['btr_acc'])
    assert 1 + tr.nfe == int(g['nfe_b'])     # adjoint.py:67 evaluates func once more per interval


@pytest.mark.parametrize('ode', ['constant', 'linear', 'sine'])
@pytest.mark.parametrize('direction', ['fwd', 'rev'])
def test_generic_problems_f64(golden, ode, direction):
    g = golden('generic_f64')
    key = '%s_%s' % (ode, direction)
    y0, t = torch.from_numpy(g[key + '.y0']), torch.from_numpy(g[key + '.t'])
    f = make_problem(ode, g.get(key + '.A'))
    tr = dopri5_port.Trace()
    out = dopri5_port.dopri5_solve(f, y0, t, 1e-7, 1e-9, trace=tr)
    np.testing.assert_allclose(out.numpy(), g[key + '.out'], rtol=1e-10, atol=1e-12)
    assert [s[2] for s in tr.steps] == list(g[key + '.tr_acc'])
    # the reference's own accuracy bar (odeint_tests.py:54-67): rel error < 1e-4... they use |diff| norm
    exact = torch.from_numpy(g[key + '.exact']).reshape(out.shape)
    assert float((out - exact).abs().max() / exact.abs().max()) < 1e-4


def make_problem(ode, A=None):
    """The analytic problems of the reference's torchdiffeq/tests/problems.py:7-57, restated."""
    if ode == 'constant':
        a, b = torch.tensor(0.2, dtype=torch.float64), torch.tensor(3.0, dtype=torch.float64)
        return lambda t, y: a + (y - (a * t + b)) ** 5
    if ode == 'sine':
        return lambda t, y: 2 * y / t + t ** 4 * torch.sin(2 * t) - t ** 2 + 4 * t ** 3
    A = torch.as_tensor(A)
    return lambda t, y: torch.mm(A.to(y), y.reshape(-1, 1)).reshape(-1)
