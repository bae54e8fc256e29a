"""GPU: the dopri5 / adjoint cases of the reference's own test files, restated against the drop-in `torchdiffeq` package
(the import the reference's model.py binds to). The reference runs them with float64 as the default dtype on
cuda:0 (torchdiffeq/tests/odeint_tests.py:7-9); so do these. Every test cites the case it restates; the bounds are the
reference's own. The fixed-grid and Adams cases of those files are out of scope (dopri5 only)."""
import math

import numpy as np
import pytest
import scipy.linalg
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
ERROR_TOL = 1e-4          # odeint_tests.py:6
EPS = 1e-12               # gradient_tests.py:7, api_tests.py:7


@pytest.fixture(autouse=True)
def default_f64(native_lib):
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(prev)


# ---- the analytic problems (torchdiffeq/tests/problems.py:7-57), as nn.Modules with the same parameters ------------------
class Constant(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Parameter(torch.tensor(0.2, device=DEV))
        self.b = torch.nn.Parameter(torch.tensor(3.0, device=DEV))

    def forward(self, t, y):
        return self.a + (y - (self.a * t + self.b)) ** 5

    def y_exact(self, t):
        return self.a * t + self.b


class Sine(torch.nn.Module):
    def forward(self, t, y):
        return 2 * y / t + t ** 4 * torch.sin(2 * t) - t ** 2 + 4 * t ** 3

    def y_exact(self, t):
        return (-0.5 * t ** 4 * torch.cos(2 * t) + 0.5 * t ** 3 * torch.sin(2 * t) + 0.25 * t ** 2 * torch.cos(2 * t) - t ** 3
                + 2 * t ** 4 + (math.pi - 0.25) * t ** 2)


class Linear(torch.nn.Module):
    def __init__(self, dim=10):
        super().__init__()
        self.dim = dim
        U = torch.randn(dim, dim, generator=torch.Generator().manual_seed(5)).to(DEV) * 0.1
        self.A = torch.nn.Parameter(2 * U - (U + U.t()))

    def forward(self, t, y):
        return torch.mm(self.A, y.reshape(self.dim, 1)).reshape(-1)

    def y_exact(self, t):
        A = self.A.detach().cpu().numpy()
        rows = [scipy.linalg.expm(A * float(ti)) @ np.ones((self.dim, 1)) for ti in t.detach().cpu()]
        return torch.tensor(np.stack(rows)).reshape(len(t), self.dim).to(DEV)


PROBLEMS = {'constant': Constant, 'linear': Linear, 'sine': Sine}


def construct_problem(ode='constant', reverse=False, npts=10):
    """problems.py:63-79: ten points on [1, 8], the exact solution, y0 = the exact solution at the first point."""
    f = PROBLEMS[ode]().to(DEV)
    t = torch.linspace(1, 8, npts).to(DEV).requires_grad_(True)
    sol = f.y_exact(t)
    if reverse:
        t = t.flip(0).clone().detach()
        sol = sol.flip(0).clone().detach()
    return f, sol[0].detach(), t, sol


def max_abs(x):
    return float(x.abs().max())


def rel_error(true, est):
    return max_abs((true - est) / true)


# ---- odeint_tests.py ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('ode', ['constant', 'linear', 'sine'])
@pytest.mark.parametrize('reverse', [False, True])
def test_dopri5_solver_error(ode, reverse):
    """odeint_tests.py:54-60 (TestSolverError.test_dopri5) and :104-110 (TestSolverBackwardsInTimeError.test_dopri5)."""
    import torchdiffeq
    f, y0, t, sol = construct_problem(ode, reverse)
    with torch.no_grad():
        y = torchdiffeq.odeint(f, y0, t, method='dopri5')
    assert rel_error(sol, y) < ERROR_TOL


@pytest.mark.parametrize('ode', ['constant', 'linear', 'sine'])
@pytest.mark.parametrize('reverse', [False, True])
def test_adjoint_solver_error(ode, reverse):
    """odeint_tests.py:62-67, :112-118 (test_adjoint: odeint_adjoint with dopri5, graph attached)."""
    import torchdiffeq
    f, y0, t, sol = construct_problem(ode, reverse)
    y = torchdiffeq.odeint_adjoint(f, y0, t, method='dopri5')
    assert rel_error(sol, y) < ERROR_TOL


def test_no_integration():
    """odeint_tests.py:147-151 (TestNoIntegration.test_dopri5): one time point returns y0."""
    import torchdiffeq
    f, y0, t, sol = construct_problem('constant', reverse=True)
    y = torchdiffeq.odeint(f, y0, t[0:1], method='dopri5')
    assert max_abs(sol[0] - y) < ERROR_TOL


# ---- gradient_tests.py ----------------------------------------------------------------------------------------------------
def test_gradcheck_dopri5():
    """gradient_tests.py:33-37 (TestGradient.test_dopri5): numerical vs analytic Jacobian with respect to y0 and the time points."""
    import torchdiffeq
    f, y0, t, _ = construct_problem()
    y0 = y0.clone().requires_grad_(True)
    func = lambda y0_, t_: torchdiffeq.odeint(f, y0_, t_, method='dopri5')
    assert torch.autograd.gradcheck(func, (y0, t))


def test_adjoint_gradients_against_odeint():
    """gradient_tests.py:45-76 (TestGradient.test_adjoint): gradients with respect to the time points and the parameters through
    odeint and through odeint_adjoint agree to 1e-12 on the (polynomial, hence exactly integrated) constant problem."""
    import torchdiffeq
    f, y0, t, _ = construct_problem()
    ys = torchdiffeq.odeint(f, y0, t, method='dopri5')
    gradys = torch.rand(ys.shape, generator=torch.Generator().manual_seed(0)).to(DEV)
    ys.backward(gradys)
    reg = (t.grad.clone(), f.a.grad.clone(), f.b.grad.clone())
    f, y0, t, _ = construct_problem()
    ys = torchdiffeq.odeint_adjoint(f, y0, t, method='dopri5')
    ys.backward(gradys)
    assert max_abs(reg[0] - t.grad) < EPS
    assert max_abs(reg[1] - f.a.grad) < EPS
    assert max_abs(reg[2] - f.b.grad) < EPS


class Cubic(torch.nn.Module):
    """gradient_tests.py:83-92: a stiff-ish spiral with a parameter matrix and a module the dynamics never use."""

    def __init__(self):
        super().__init__()
        self.A = torch.nn.Parameter(torch.tensor([[-0.1, 2.0], [-2.0, -0.1]]))
        self.unused_module = torch.nn.Linear(2, 5)

    def forward(self, t, y):
        return torch.mm(y ** 3, self.A)


def cubic_problem():
    y0 = torch.tensor([[2., 0.]]).to(DEV).requires_grad_(True)
    t = torch.linspace(0., 25., 10).to(DEV).requires_grad_(True)
    return Cubic().to(DEV), y0, t


def test_dopri5_adjoint_against_dopri5():
    """gradient_tests.py:98-116 (TestCompareAdjointGradient.test_dopri5_adjoint_against_dopri5), bounds 3e-4 / 1e-4 / 2e-3; the unused
    module's parameters get exact zeros, not None."""
    import torchdiffeq
    func, y0, t = cubic_problem()
    ys = torchdiffeq.odeint_adjoint(func, y0, t, method='dopri5')
    gradys = torch.rand(ys.shape, generator=torch.Generator().manual_seed(1)).to(DEV) * 0.1
    ys.backward(gradys)
    adj = (y0.grad.clone(), t.grad.clone(), func.A.grad.clone())
    assert max_abs(func.unused_module.weight.grad) == 0
    assert max_abs(func.unused_module.bias.grad) == 0
    func, y0, t = cubic_problem()
    ys = torchdiffeq.odeint(func, y0, t, method='dopri5')
    ys.backward(gradys)
    assert max_abs(y0.grad - adj[0]) < 3e-4
    assert max_abs(t.grad - adj[1]) < 1e-4
    assert max_abs(func.A.grad - adj[2]) < 2e-3


# ---- api_tests.py ---------------------------------------------------------------------------------------------------------
def test_tuple_state_dopri5():
    """api_tests.py:19-28 (TestCollectionState.test_dopri5): a tuple state through a plain callable; both members within 1e-12
    of the exact solution (the reference compares signed differences)."""
    import torchdiffeq
    f, y0, t, sol = construct_problem()
    tuple_f = lambda t_, y: (f(t_, y[0]), f(t_, y[1]))
    with torch.no_grad():
        ys = torchdiffeq.odeint(tuple_f, (y0, y0), t, method='dopri5')
    assert float((sol - ys[0]).max()) < EPS
    assert float((sol - ys[1]).max()) < EPS


@pytest.mark.parametrize('member', [0, 1])
def test_tuple_state_dopri5_gradient(member):
    """api_tests.py:30-38 (test_dopri5_gradient): gradcheck through a tuple state and a lambda that closes over the module."""
    import torchdiffeq
    f, y0, t, _ = construct_problem()
    y0 = y0.clone().requires_grad_(True)
    tuple_f = lambda t_, y: (f(t_, y[0]), f(t_, y[1]))
    func = lambda y0_, t_: torchdiffeq.odeint(tuple_f, (y0_, y0_), t_, method='dopri5')[member]
    assert torch.autograd.gradcheck(func, (y0, t))
