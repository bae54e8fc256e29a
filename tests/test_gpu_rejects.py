"""GPU: rejected steps on the fused route. At random initialisation the forward solve from t = 0 accepts every step
(SURVEY appendix B), so the accept/reject machinery of the step kernels - ping-pong of the current state buffer, FSAL reuse
of k1 after a reject, the stale interpolant invariant of dopri5.py:121 - is exercised here by
integrating BACKWARDS from the reference's own y(1): the oracle (and the unmodified reference, tools/make_golden.py) reject
one to three attempts on that span for every feature-map shape. Every case asserts that the oracle really rejected."""
import pytest
import torch

from conftest import load_odefunc, odefunc_params
from oracle import dopri5_port, odefunc_port

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL_OUT = 1e-4
SHAPES = {'cifar_res_n8': '8x8', 'mnist_conv_n9': '6x6', 'mnist_res_n5': '7x7', 'cifar_oneshot_n3': '16x16', 'mnist_oneshot_n3': '14x14'}


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def check_against_oracle(func, p, y, t, tol=1e-3):
    from node_b200 import odeint, solver
    tr = dopri5_port.Trace()
    with torch.no_grad():
        ref = dopri5_port.dopri5_solve(lambda a, b: odefunc_port.odefunc_forward(p, a, b), y, t, tol, tol, trace=tr)
        out = odeint(func, y.to(DEV), t.to(DEV), rtol=tol, atol=tol, method='dopri5')
    st = dict(solver.last_stats)
    oracle_acc = [bool(s[2]) for s in tr.steps]
    assert not all(oracle_acc), 'this case is meant to contain rejected steps'
    assert st['route'] == 'fused'
    assert st['nfe'] == tr.nfe == 2 + 6 * len(tr.steps)
    assert [bool(a) for a in st['trace']['accepted']] == oracle_acc
    assert st['n_reject'] == oracle_acc.count(False)
    dts = torch.tensor([s[1] for s in tr.steps], dtype=torch.float64)
    got = torch.tensor(st['trace']['dt'], dtype=torch.float64)
    err_dt = (got - dts).abs() / dts.abs()
    # dt trace within 1e-5 relative. The FIRST dt gets 5e-5: it is (0.01 / d2)^(1/5) with d2 the difference
    # quotient |f(t0 + h0, y0 + h0 f0) - f0| / h0 (misc.py:133-141), which amplifies the ~1e-6 error of an fp32 evaluation by
    # |f| / |f1 - f0| ~ 1e2 on these spans; every later dt is a factor ratio^(1/10) of its predecessor and self-corrects.
    assert float(err_dt[0]) < 5e-5 and float(err_dt[1:].max()) < 1e-5, err_dt
    t0s = torch.tensor([s[0] for s in tr.steps], dtype=torch.float64)
    assert float((torch.tensor(st['trace']['t'], dtype=torch.float64) - t0s).abs().max()) < 1e-5
    for i in range(len(t)):                                                   # every output time, not only the last
        assert rel(out[i].cpu(), ref[i]) < TOL_OUT, i
    return oracle_acc


@pytest.mark.parametrize('engine', ['dense', 'strip'])
@pytest.mark.parametrize('name', sorted(SHAPES))
def test_fused_rejected_steps_every_shape(native_lib, golden, monkeypatch, name, engine):
    if engine == 'strip':
        if SHAPES[name] != '8x8':
            pytest.skip('only 8x8 has two step engines')
        monkeypatch.setenv('NODE_B200_STEP8', '0')
    g = golden(name)
    acc = check_against_oracle(load_odefunc(g, DEV), odefunc_params(g), torch.from_numpy(g['out'][-1]), torch.tensor([1.0, 0.0]))
    assert acc.count(False) >= 1


@pytest.mark.parametrize('name', ['cifar_res_n8', 'mnist_conv_n9', 'cifar_oneshot_n3'])
def test_fused_rejected_steps_with_dense_output(native_lib, golden, name):
    """Output times inside accepted steps that are separated by rejected attempts: the interpolant must stay the one of
    the last ACCEPTED step while the controller retries (dopri5.py:92,121)."""
    g = golden(name)
    check_against_oracle(load_odefunc(g, DEV), odefunc_params(g), torch.from_numpy(g['out'][-1]),
                         torch.tensor([1.0, 0.9, 0.75, 0.5, 0.3, 0.25, 0.0]))


@pytest.mark.parametrize('engine', ['dense', 'strip'])
@pytest.mark.parametrize('n', [452, 1190])
def test_large_batch_step_kernels_reject(native_lib, golden, monkeypatch, n, engine):
    """The variants the benchmark runs (k_step8 on CTA pairs over several rounds; k_step<8,8,2> above 444 images), ragged
    tails included (452 = 113 super-tiles: the last pair has no peer tile), on a span with rejected steps."""
    if engine == 'strip':
        monkeypatch.setenv('NODE_B200_STEP8', '0')
    g = golden('cifar_res_n8')
    base = torch.from_numpy(g['out'][-1])
    gen = torch.Generator().manual_seed(n)
    reps = (n + base.shape[0] - 1) // base.shape[0]
    y = (base.repeat(reps, 1, 1, 1)[:n] * (1.0 + 0.05 * torch.randn(n, 1, 1, 1, generator=gen))).contiguous()
    check_against_oracle(load_odefunc(g, DEV), odefunc_params(g), y, torch.tensor([1.0, 0.0]))
