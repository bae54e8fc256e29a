"""Pins the oracle against the UNMODIFIED reference and writes tests/golden/*.npz.

Runs only in the build container (it imports /root/reference, which does not exist on the GPU
box):   python tools/make_golden.py
  1. imports the reference's model.py + its pinned torchdiffeq from /root/reference;
  2. checks oracle/odefunc_port.py and oracle/dopri5_port.py against it BIT FOR BIT on CPU
     (state, dt trace, accept/reject sequence, NFE, adjoint gradients) - any mismatch aborts;
  3. checks that node_b200.models draws the same initial weights / state_dict keys;
  4. records inputs, weights and the reference's outputs as the committed golden vectors.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
sys.path[:0] = [os.path.join(REF, 'torchdiffeq'), os.path.join(REF, 'expman'), REF]
import model as ref_model                                   # noqa: E402  the reference, as shipped
import torchdiffeq as ref_tde                               # noqa: E402
from torchdiffeq._impl import dopri5 as ref_dopri5          # noqa: E402
sys.path.insert(0, ROOT)
from oracle import dopri5_port, odefunc_port                # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(GOLD, exist_ok=True)
torch.set_num_threads(8)


class RefTrace(object):
    """Records the reference's attempted steps by wrapping Dopri5Solver._adaptive_dopri5_step."""

    def __enter__(self):
        self.steps = []
        self.orig = ref_dopri5.Dopri5Solver._adaptive_dopri5_step
        outer = self

        def wrapped(solver, rk_state):
            new = outer.orig(solver, rk_state)
            accepted = bool(new.t1 > rk_state.t1)
            outer.steps.append((float(rk_state.t1), float(rk_state.dt), accepted))
            return new
        ref_dopri5.Dopri5Solver._adaptive_dopri5_step = wrapped
        return self

    def __exit__(self, *a):
        ref_dopri5.Dopri5Solver._adaptive_dopri5_step = self.orig


def same(a, b, what):
    assert a.shape == b.shape and bool((a == b).all()), '%s: oracle differs from the reference (max |d| = %g)' % (
        what, float((a - b).abs().max()))


def trace_arrays(steps):
    return (np.array([s[0] for s in steps]), np.array([s[1] for s in steps]), np.array([s[2] for s in steps], dtype=bool))


def odenet_case(name, in_ch, size, downsample, N, tol=1e-3, t1=1, store_full=True, seed=0):
    torch.manual_seed(seed)
    net = ref_model.ODENet(in_ch, n_filters=64, downsample=downsample, tol=tol, t1=t1).eval()
    x = torch.rand(N, in_ch, size, size)
    func = net.odeblock.odefunc
    p = {k: v.detach() for k, v in odefunc_port.params_from_module(func).items()}
    with torch.no_grad():
        h0 = net.downsample(x)
        # dynamics: restatement == reference module, bit for bit
        tt = torch.tensor(0.37)
        same(odefunc_port.odefunc_forward(p, tt, h0), func(tt, h0), name + ' odefunc')
        fold = odefunc_port.odefunc_forward_folded(p, tt, h0)
        f_ref = func(tt, h0)
        assert float((fold - f_ref).abs().max()) < 5e-6 * float(f_ref.abs().max() + 1), 'time-map fold drifted'
        func.nfe = 0
        t = net.odeblock.integration_time
        with RefTrace() as rt:
            ref_out = ref_tde.odeint(func, h0, t, rtol=tol, atol=tol, method='dopri5')
        nfe = func.nfe
        tr = dopri5_port.Trace()
        ora_out = dopri5_port.dopri5_solve(lambda a, b: odefunc_port.odefunc_forward(p, a, b), h0, t, tol, tol, trace=tr)
        same(ora_out, ref_out, name + ' solve')
        assert tr.nfe == nfe, (tr.nfe, nfe)
        assert [(s[0], s[1], s[2]) for s in tr.steps] == rt.steps, name + ' step trace'
        func.nfe = 0
        logits = net(x)
    ts, dts, acc = trace_arrays(rt.steps)
    rec = dict(seed=seed, N=N, in_ch=in_ch, size=size, tol=tol, t=t.numpy(), nfe=nfe, tr_t=ts, tr_dt=dts, tr_acc=acc,
               tr_ratio=np.array([max(s[3]) for s in tr.steps]), logits=logits.numpy(),
               out_mean=float(ref_out[-1].mean()), out_std=float(ref_out[-1].std()),
               out_slice=ref_out[-1][0, 0, 0, :4].numpy(), f037_slice=f_ref[0, :2, 0, :4].numpy())
    if store_full:
        rec.update(h0=h0.numpy(), out=ref_out.numpy(), f037=f_ref.numpy())
        rec.update({'p.' + k: v.numpy() for k, v in p.items()})
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **rec)
    print('%-28s N=%-4d state=%s nfe=%d steps=%d rejects=%d' % (name, N, tuple(h0.shape[1:]), nfe, len(ts), int((~acc).sum())))
    return net, x


def mirror_case():
    """node_b200.models draws the same weights under the same seed and has the same keys."""
    sys.path.insert(0, os.path.join(ROOT, 'neural-ode-features_b200'))
    from node_b200 import models
    for ds, in_ch in (('residual', 3), ('convolution', 1), ('one-shot', 3), ('minimal', 1), ('ode', 3), ('ode2', 3)):
        torch.manual_seed(3)
        a = ref_model.ODENet(in_ch, n_filters=64, downsample=ds, dropout=0.5, adjoint=True, t1=[.25, 1]).state_dict()
        torch.manual_seed(3)
        b = models.ODENet(in_ch, n_filters=64, downsample=ds, dropout=0.5, adjoint=True, t1=[.25, 1]).state_dict()
        assert list(a.keys()) == list(b.keys()), ds
        for k in a:
            assert bool((a[k] == b[k]).all()), (ds, k)
    print('models mirror: state_dict keys and seeded initial weights identical for 6 downsamplers')


def adjoint_case(name, N, tol=1e-3, seed=0, in_ch=3, size=32, downsample='residual', n_filters=64):
    torch.manual_seed(seed)
    net = ref_model.ODENet(in_ch, n_filters=n_filters, downsample=downsample, tol=tol, adjoint=True).train()
    x = torch.rand(N, in_ch, size, size)
    y = torch.randint(0, 10, (N,))
    func = net.odeblock.odefunc
    h0 = net.downsample(x).detach().requires_grad_(True)
    t = net.odeblock.integration_time
    func.nfe = 0
    with RefTrace() as rt:
        out = ref_tde.odeint_adjoint(func, h0, t, rtol=tol, atol=tol, method='dopri5')
        nfe_f = func.nfe
        func.nfe = 0
        n_fwd = len(rt.steps)
        loss = torch.nn.functional.cross_entropy(net.classifier(out[-1]), y)
        loss.backward()
    nfe_b = func.nfe
    g_out = torch.autograd.grad(torch.nn.functional.cross_entropy(net.classifier(out[-1].detach().requires_grad_(True)), y),
                                [], allow_unused=True) if False else None
    params = list(func.parameters())
    flat_grad = torch.cat([q.grad.reshape(-1) for q in params])
    # oracle adjoint, autograd VJP and hand-derived VJP
    p = {k: v.detach() for k, v in odefunc_port.params_from_module(func).items()}
    o = out.detach()
    go = torch.zeros_like(o)
    tmp = o[-1].clone().requires_grad_(True)
    torch.nn.functional.cross_entropy(net.classifier(tmp), y).backward()
    go[-1] = tmp.grad
    for q in params:
        q.grad = None
    tr = dopri5_port.Trace()
    gy, gt, gp = dopri5_port.adjoint_backward(lambda a, b: func(a, b), params, t, o, go, tol, tol, trace=tr)
    same(gy, h0.grad, name + ' adjoint grad_y0')
    same(gp, flat_grad, name + ' adjoint grad_params')
    bsteps = rt.steps[n_fwd:]
    assert [(s[0], s[1], s[2]) for s in tr.steps] == bsteps, name + ' backward trace'
    gy2, gt2, gp2 = dopri5_port.adjoint_backward(
        lambda a, b: odefunc_port.odefunc_forward(p, a, b), params, t, o, go, tol, tol,
        vjp=lambda a, b, c: odefunc_port.odefunc_vjp(p, a, b, c))
    e1 = float((gy2 - gy).abs().max() / gy.abs().max())
    e2 = float((gp2 - gp).abs().max() / gp.abs().max())
    tr2 = dopri5_port.Trace()
    dopri5_port.adjoint_backward(lambda a, b: odefunc_port.odefunc_forward(p, a, b), params, t, o, go, tol, tol,
                                 vjp=lambda a, b, c: odefunc_port.odefunc_vjp(p, a, b, c), trace=tr2)
    same_seq = [s[2] for s in tr2.steps] == [s[2] for s in tr.steps]
    # An adaptive solve at tol 1e-3 is only reproducible to ~1e-5 when the accept/reject sequence is the same: when a ratio
    # sits next to 1 a 1e-6 perturbation (hand-derived VJP vs autograd) flips a decision and the two discretisations differ
    # by the solver's own error (1e-2). Record which it is; the per-evaluation VJP check (1e-5) is separate.
    if not (e1 < 1e-4 and e2 < 1e-4):
        # per evaluation the hand-derived VJP is within 1e-6 of autograd (checked for every case by tests/test_oracle.py);
        # what is left is the conditioning of the reverse-time solve itself, recorded below against float64.
        print('   NOTE: 1e-6 per-evaluation VJP differences grow to y %.1e p %.1e over this reverse solve' % (e1, e2))
    if not same_seq:
        print('   NOTE: hand-derived VJP changes the step sequence of this case: %s vs %s (min |ratio - 1| = %.1e)' % (
            [int(s[2]) for s in tr2.steps], [int(s[2]) for s in tr.steps], min(abs(max(s[3]) - 1) for s in tr.steps)))
    ts, dts, acc = trace_arrays(bsteps)
    # Conditioning of the reverse solve, measured on the reference's own arithmetic: perturb y(t1) by 3e-6 relative (the size of
    # the disagreement between two correct fp32 forward solves on different hardware / summation orders) and see how far
    # the gradient moves. tools/adjoint_sensitivity.py scans the amplitude: 1e-6 moves it by ~1e-6, 3e-6 by up to 1e-2.
    sens_y, sens_p = 0.0, 0.0
    for draw in range(2):
        gen = torch.Generator().manual_seed(100 + draw)
        o2 = o.clone()
        o2[-1] = o2[-1] * (1 + 3e-6 * torch.randn(o2[-1].shape, generator=gen))
        gys, _, gps = dopri5_port.adjoint_backward(lambda a, b: func(a, b), params, t, o2, go, tol, tol)
        sens_y = max(sens_y, float((gys - gy).abs().max() / gy.abs().max()))
        sens_p = max(sens_p, float((gps - gp).abs().max() / gp.abs().max()))
    print('   reference adjoint under a 3e-6 perturbation of y(t1): grad_y0 moves %.2e  grad_params %.2e' % (sens_y, sens_p))
    # The same adjoint in float64 (the reference itself, every tensor widened): how far the reference's own fp32 gradient
    # is from the exact one - gradients of a ReLU network are discontinuous in the state, so this, not 1e-7, is the scale
    # against which an independent fp32 implementation can be judged.
    import copy
    f64 = copy.deepcopy(func).double()
    f64.nfe = 0
    h64 = h0.detach().double().requires_grad_(True)
    cls64 = copy.deepcopy(net.classifier).double()
    out64 = ref_tde.odeint_adjoint(f64, h64, t.double(), rtol=tol, atol=tol, method='dopri5')
    torch.nn.functional.cross_entropy(cls64(out64[-1]), y).backward()
    gy64 = h64.grad
    gp64 = torch.cat([q.grad.reshape(-1) for q in f64.parameters()])
    r1 = float((h0.grad.double() - gy64).abs().max() / gy64.abs().max())
    r2 = float((flat_grad.double() - gp64).abs().max() / gp64.abs().max())
    print('   reference fp32 adjoint vs its float64 self: grad_y0 %.2e  grad_params %.2e (max-norm relative)' % (r1, r2))
    rec = dict(seed=seed, N=N, tol=tol, t=t.numpy(), h0=h0.detach().numpy(), out=o.numpy(), grad_out=go.numpy(),
               grad_y0=h0.grad.numpy(), grad_params=flat_grad.numpy(), grad_t=gt.numpy(), nfe_f=nfe_f, nfe_b=nfe_b,
               btr_t=ts, btr_dt=dts, btr_acc=acc, labels=y.numpy(),
               grad_y0_f64=gy64.float().numpy(), grad_params_f64=gp64.float().numpy(),    # stored in float32: 6e-8 is exact enough
               ref_err_y0=r1, ref_err_params=r2,
               btr_ratio=np.array([max(s[3]) for s in tr.steps]), hand_vjp_same_sequence=same_seq,
               sens_y0=sens_y, sens_params=sens_p, hand_vjp_dev_y0=e1, hand_vjp_dev_params=e2)
    rec.update({'p.' + k: v.numpy() for k, v in p.items()})
    if n_filters > 64:                       # the wide fixtures keep the scalars (ref_err_*), not the float64 gradient arrays
        rec.pop('grad_y0_f64'), rec.pop('grad_params_f64')
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **rec)
    print('%-28s N=%-4d nfe_f=%d nfe_b=%d bwd steps=%d rejects=%d  hand-VJP rel err y %.1e p %.1e' % (
        name, N, nfe_f, nfe_b, len(ts), int((~acc).sum()), e1, e2))


def generic_cases():
    """Analytic problems of the reference's own test-suite (torchdiffeq/tests/problems.py), float64:
    the reference's odeint outputs become golden vectors for the generic-callable route."""
    sys.path.insert(0, os.path.join(REF, 'torchdiffeq', 'tests'))
    import problems
    torch.manual_seed(0)
    rec = {}
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        for ode in ('constant', 'linear', 'sine'):
            for reverse in (False, True):
                torch.manual_seed(1)
                f, y0, t, sol = problems.construct_problem('cpu', ode=ode, reverse=reverse)
                with RefTrace() as rt:
                    ref = ref_tde.odeint(f, y0, t.detach(), method='dopri5')
                tr = dopri5_port.Trace()
                ora = dopri5_port.dopri5_solve(f, y0, t.detach(), 1e-7, 1e-9, trace=tr)
                same(ora.detach(), ref.detach(), 'generic %s rev=%s' % (ode, reverse))
                assert [(s[0], s[1], s[2]) for s in tr.steps] == rt.steps
                key = '%s_%s' % (ode, 'rev' if reverse else 'fwd')
                rec[key + '.y0'] = y0.detach().numpy()
                rec[key + '.t'] = t.detach().numpy()
                rec[key + '.out'] = ref.detach().numpy()
                rec[key + '.exact'] = sol.detach().numpy()
                rec[key + '.tr_dt'] = np.array([s[1] for s in rt.steps])
                rec[key + '.tr_acc'] = np.array([s[2] for s in rt.steps], dtype=bool)
                if ode == 'linear':
                    rec[key + '.A'] = f.A.detach().numpy()
                print('generic %-9s reverse=%-5s steps=%d rejects=%d' % (ode, reverse, len(rt.steps), sum(1 for s in rt.steps if not s[2])))
        # tuple state (api_tests.py:19-38 flavour)
        torch.manual_seed(2)
        f, y0, t, _ = problems.construct_problem('cpu', ode='linear')
        ft = lambda tt, y: (f(tt, y[0]), -0.5 * y[1])
        z0 = torch.randn(3, 4)
        ref = ref_tde.odeint(ft, (y0, z0), t.detach(), method='dopri5')
        ora = dopri5_port.dopri5_solve(ft, (y0, z0), t.detach(), 1e-7, 1e-9)
        same(ora[0].detach(), ref[0].detach(), 'tuple 0')
        same(ora[1].detach(), ref[1].detach(), 'tuple 1')
        rec.update({'tuple.A': f.A.detach().numpy(), 'tuple.y0': y0.numpy(), 'tuple.z0': z0.numpy(), 'tuple.t': t.detach().numpy(),
                    'tuple.out0': ref[0].detach().numpy(), 'tuple.out1': ref[1].detach().numpy()})
    finally:
        torch.set_default_dtype(prev)
    np.savez_compressed(os.path.join(GOLD, 'generic_f64.npz'), **rec)


def unrolled_case(name, N, tol=1e-3, seed=0, in_ch=3, size=32, downsample='residual', times=None):
    """gradients of the reference's NON-adjoint odeint (autograd unrolled through the solver, controller included;
    model.py:359 with adjoint=False) for the ODE-Net dynamics, in float32 and float64, and the check that node_b200.unrolled's
    solver loop (plain torch ops around the dynamics) reproduces the reference's unrolled gradient BIT FOR BIT on CPU when it is
    handed the same eager dynamics."""
    import copy
    sys.path.insert(0, os.path.join(ROOT, 'neural-ode-features_b200'))
    from node_b200 import unrolled
    torch.manual_seed(seed)
    net = ref_model.ODENet(in_ch, n_filters=64, downsample=downsample, tol=tol, adjoint=False).train()
    x = torch.rand(N, in_ch, size, size)
    func = net.odeblock.odefunc
    h0 = net.downsample(x).detach().requires_grad_(True)
    t = (net.odeblock.integration_time if times is None else torch.tensor(times)).clone().requires_grad_(True)
    func.nfe = 0
    with RefTrace() as rt:
        out = ref_tde.odeint(func, h0, t, rtol=tol, atol=tol, method='dopri5')
    nfe = func.nfe
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(seed + 1))
    out.backward(go)
    params = list(func.parameters())
    gp = torch.cat([q.grad.reshape(-1) for q in params])
    gy, gt = h0.grad.clone(), t.grad.clone()
    # node_b200.unrolled with the same eager dynamics
    for q in params:
        q.grad = None
    h1 = h0.detach().clone().requires_grad_(True)
    t1 = t.detach().clone().requires_grad_(True)
    st = {}
    mine = unrolled.solve(lambda tt, yy: (func(tt, yy[0]),), (h1,), t1, tol, tol, {}, stats=st)[0]
    same(mine.detach(), out.detach(), name + ' unrolled forward')
    mine.backward(go)
    same(h1.grad, gy, name + ' unrolled grad_y0')
    same(t1.grad, gt, name + ' unrolled grad_t')
    same(torch.cat([q.grad.reshape(-1) for q in params]), gp, name + ' unrolled grad_params')
    assert st['nfe'] == nfe and (st['n_accept'], st['n_reject']) == (sum(s[2] for s in rt.steps), sum(not s[2] for s in rt.steps))
    # float64 self and the adjoint's gradient for scale
    f64 = copy.deepcopy(func).double()
    h64 = h0.detach().double().requires_grad_(True)
    t64 = t.detach().double().requires_grad_(True)
    o64 = ref_tde.odeint(f64, h64, t64, rtol=tol, atol=tol, method='dopri5')
    o64.backward(go.double())
    gp64 = torch.cat([q.grad.reshape(-1) for q in f64.parameters()])
    r1 = float((gy.double() - h64.grad).abs().max() / h64.grad.abs().max())
    r2 = float((gp.double() - gp64).abs().max() / gp64.abs().max())
    for q in params:
        q.grad = None
    h2 = h0.detach().clone().requires_grad_(True)
    oa = ref_tde.odeint_adjoint(func, h2, t.detach(), rtol=tol, atol=tol, method='dopri5')
    oa.backward(go)
    gpa = torch.cat([q.grad.reshape(-1) for q in params])
    a1 = float((h2.grad - gy).abs().max() / gy.abs().max())
    a2 = float((gpa - gp).abs().max() / gp.abs().max())
    p = {k: v.detach() for k, v in odefunc_port.params_from_module(func).items()}
    ts, dts, acc = trace_arrays(rt.steps)
    rec = dict(seed=seed, N=N, tol=tol, t=t.detach().numpy(), h0=h0.detach().numpy(), out=out.detach().numpy(), grad_out=go.numpy(),
               grad_y0=gy.numpy(), grad_params=gp.numpy(), grad_t=gt.numpy(), nfe=nfe, tr_t=ts, tr_dt=dts, tr_acc=acc,
               grad_y0_f64=h64.grad.float().numpy(), grad_params_f64=gp64.float().numpy(), grad_t_f64=t64.grad.float().numpy(),
               ref_err_y0=r1, ref_err_params=r2, adjoint_dev_y0=a1, adjoint_dev_params=a2)
    rec.update({'p.' + k: v.numpy() for k, v in p.items()})
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **rec)
    print('%-28s N=%-3d nfe=%d steps=%d rejects=%d | fp32 vs fp64 self: y %.1e p %.1e | adjoint vs unrolled: y %.1e p %.1e' % (
        name, N, nfe, len(ts), int((~acc).sum()), r1, r2, a1, a2))


def unrolled_cases():
    unrolled_case('unrolled_cifar_n4', 4)
    unrolled_case('unrolled_cifar_n3_t3', 3, times=[0.0, 0.4, 1.0], seed=3)
    unrolled_case('unrolled_cifar_rev_n2', 2, times=[1.0, 0.0], seed=4)
    unrolled_case('unrolled_mnist_conv_n3', 3, in_ch=1, size=28, downsample='convolution', seed=1)     # 6x6
    unrolled_case('unrolled_cifar_oneshot_n2', 2, downsample='one-shot', seed=2, tol=1e-2)              # 16x16


def adjoint_cases():
    adjoint_case('adjoint_cifar_n4', 4)
    adjoint_case('adjoint_mnist_conv_n3', 3, in_ch=1, size=28, downsample='convolution')     # 6x6
    adjoint_case('adjoint_mnist_res_n3', 3, in_ch=1, size=28, downsample='residual')         # 7x7
    adjoint_case('adjoint_cifar_oneshot_n2', 2, downsample='one-shot')                       # 16x16
    adjoint_case('adjoint_mnist_oneshot_n2', 2, in_ch=1, size=28, downsample='one-shot')     # 14x14
    adjoint_case('adjoint_cifar_n32', 32, seed=5)                                            # a larger batch (averaging over images)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'adjoint':
        adjoint_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'wide':
        adjoint_case('adjoint_cifar_c128_n2', 2, n_filters=128, seed=7)       # wide dynamics (reproduce.sh:21-25 trains 256 filters with --adjoint)
        adjoint_case('adjoint_cifar_c256_n2', 2, n_filters=256, seed=8)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'unrolled':
        unrolled_cases()
        sys.exit(0)
    mirror_case()
    generic_cases()
    odenet_case('cifar_res_n8', 3, 32, 'residual', 8)
    odenet_case('cifar_res_n8_t10', 3, 32, 'residual', 4, t1=np.linspace(0, 1, 10).tolist())
    odenet_case('cifar_res_n7_tol1e-4', 3, 32, 'residual', 7, tol=1e-4)
    odenet_case('mnist_conv_n9', 1, 28, 'convolution', 9)        # 6x6 state, 3 images per tile + ragged tail
    odenet_case('mnist_res_n5', 1, 28, 'residual', 5)            # 7x7 state
    odenet_case('cifar_oneshot_n3', 3, 32, 'one-shot', 3)        # 16x16 state: one image spans two M tiles
    odenet_case('mnist_oneshot_n3', 1, 28, 'one-shot', 3)        # 14x14
    odenet_case('cifar_res_n128', 3, 32, 'residual', 128, store_full=False)   # SURVEY appendix B (seeds only)
    odenet_case('mnist_conv_n128', 1, 28, 'convolution', 128, store_full=False)  # BASELINE cfg1 (seeds only)
    adjoint_cases()
    adjoint_case('adjoint_cifar_c128_n2', 2, n_filters=128, seed=7)
    adjoint_case('adjoint_cifar_c256_n2', 2, n_filters=256, seed=8)
    unrolled_cases()
    print('golden vectors written to', GOLD)
