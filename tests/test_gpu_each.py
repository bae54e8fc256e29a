"""GPU: batched independent solvers (evaluate.py:109-126 runs batch_size = 1 to record per-image NFE): every sample
of odeint_each is the batch-1 solve - same result, same NFE / accept / reject counts - and against the oracle's batch-1 solves."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _func(seed=0, scale=1.0):
    from node_b200 import models
    torch.manual_seed(seed)
    f = models.ODEfunc(64).to(DEV)
    with torch.no_grad():
        for p in f.parameters():
            p.mul_(scale)
    return f


@pytest.mark.parametrize('hw,n,times,tol', [(8, 37, [0.0, 1.0], 1e-3), (8, 5, [0.0, 0.4, 1.0], 1e-4), (7, 9, [1.0, 0.0], 1e-3), (6, 3, [0.0, 1.0], 1e-2)])
def test_each_equals_batch1(native_lib, hw, n, times, tol):
    import torchdiffeq
    from node_b200 import each, solver
    f = _func(hw, 3.0)
    y0 = torch.randn(n, 64, hw, hw, device=DEV) * torch.linspace(0.2, 3.0, n, device=DEV).view(n, 1, 1, 1)
    t = torch.tensor(times, device=DEV)
    with torch.no_grad():
        f.nfe = 0
        out, stats = each.odeint_each(f, y0, t, rtol=tol, atol=tol, lanes=4)
        assert f.nfe == sum(s['nfe'] for s in stats)
        seqs = set()
        for i in range(n):
            f.nfe = 0
            ref = torchdiffeq.odeint(f, y0[i:i + 1], t, rtol=tol, atol=tol, method='dopri5')
            st = solver.last_stats
            assert (stats[i]['nfe'], stats[i]['n_accept'], stats[i]['n_reject']) == (st['nfe'], st['n_accept'], st['n_reject']) and f.nfe == st['nfe']
            assert torch.equal(out[:, i:i + 1], ref), (i, float((out[:, i:i + 1] - ref).abs().max()))
            seqs.add((st['n_accept'], st['n_reject']))
    assert len(seqs) > 1 or n < 5, 'the samples were meant to take different step sequences'


def test_each_few_steps_enqueued(native_lib):
    """A solve that needs more attempted steps than were enqueued is finished by the ordinary route."""
    import torchdiffeq
    from node_b200 import each
    f = _func(1, 3.0)
    y0 = torch.randn(6, 64, 8, 8, device=DEV) * 2.0
    t = torch.tensor([0.0, 1.0], device=DEV)
    with torch.no_grad():
        out, stats = each.odeint_each(f, y0, t, rtol=1e-4, atol=1e-4, lanes=3, steps=2)
        for i in range(6):
            ref = torchdiffeq.odeint(f, y0[i:i + 1], t, rtol=1e-4, atol=1e-4, method='dopri5')
            assert torch.equal(out[:, i:i + 1], ref)
        assert all(s['n_accept'] + s['n_reject'] > 2 for s in stats)


def test_each_matches_oracle_batch1(native_lib):
    """Per-sample step sequences against the oracle's batch-1 solves (the reference's evaluate.py nfe semantics)."""
    import copy
    from oracle import dopri5_port, odefunc_port
    from node_b200 import each
    f = _func(2, 3.0)
    y0 = torch.randn(4, 64, 8, 8, device=DEV) * torch.tensor([0.3, 1.0, 2.0, 4.0], device=DEV).view(4, 1, 1, 1)
    t = torch.tensor([0.0, 1.0], device=DEV)
    with torch.no_grad():
        out, stats = each.odeint_each(f, y0, t, rtol=1e-3, atol=1e-3)
        p = odefunc_port.params_from_module(copy.deepcopy(f).cpu())
        for i in range(4):
            tr = dopri5_port.Trace()
            ref = dopri5_port.dopri5_solve(lambda a, b: odefunc_port.odefunc_forward(p, a, b), y0[i:i + 1].cpu(), t.cpu(), 1e-3, 1e-3, trace=tr)
            assert (tr.nfe, tr.n_accept, tr.n_reject) == (stats[i]['nfe'], stats[i]['n_accept'], stats[i]['n_reject']), (i, tr.nfe, stats[i])
            err = float((out[:, i:i + 1].cpu() - ref).abs().max() / ref.abs().max())
            assert err <= 1e-4, (i, err)
