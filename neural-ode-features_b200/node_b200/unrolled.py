"""Differentiable non-adjoint `odeint`: the gradient the reference gets by letting autograd unroll every solver
operation (train.py without --adjoint, model.py:359; gradient_tests.py:19-43) - INCLUDING the terms through the step-size
controller: in the reference `dt` is a tensor in the graph (misc.py:160-170 divides it by a factor computed from the error
ratio of the state, dopri5.py:80 derives the first step from norms of y0 / f0), the stage times `t0 + alpha_i * dt` feed the
time channel of the dynamics and the output is a Hermite interpolant evaluated at (t - t0) / (t1 - t0) (interp.py:54-65). The
adjoint ODE (`odeint_adjoint`) is a different, continuous-time gradient; the two agree only to the solver tolerance.

What runs where: the solver loop below is recorded by autograd, so its Runge-Kutta combinations, error ratio and controller are
ATen elementwise ops exactly in the reference's order of operations (the arithmetic that decides accept / reject is the same as
the native kernels'); every evaluation of the recognised ODE-Net dynamics - forward AND its vector-Jacobian product in the
backward pass - is this repo's tcgen05 kernels (`_NativeDynamics`: node_b200_odefunc_forward / node_b200_odefunc_vjp). Other
callables are simply called and differentiated by autograd. Memory grows with the number of evaluations, as in the reference.
"""
import torch
import torch.nn as nn

from . import solver as _solver

_ALPHA = (1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0)                                            # dopri5.py:11-31
_BETA = ((1 / 5,), (3 / 40, 9 / 40), (44 / 45, -56 / 15, 32 / 9), (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
         (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656), (35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84))
_C_ERR = (35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720, -2187 / 6784 - -12231 / 42400,
          11 / 84 - 649 / 6300, -1.0 / 60.0)
_C_MID = (6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
          187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2)     # dopri5.py:33-36


class _NativeDynamics(torch.autograd.Function):
    """f(t, y) of the recognised ODE-Net dynamics as ONE autograd node: forward = the fused evaluation kernel, backward = the
    adjoint's VJP kernels (data gradient, parameter gradients in `func.parameters()` order, time gradient)."""

    @staticmethod
    def forward(ctx, func, t, y, *params):
        ctx.func = func
        y = y.contiguous()
        ctx.save_for_backward(t, y)
        return _solver.odefunc_forward(func, float(t), y)

    @staticmethod
    def backward(ctx, g):
        t, y = ctx.saved_tensors
        func = ctx.func
        _, vy, vt, vp = _solver.odefunc_vjp(func, t.detach().to(torch.float32), y, -g.contiguous())   # VJP kernels take cotangent -a
        grads, o = [], 0
        for p in func.parameters():
            n = p.numel()
            grads.append(vp[o:o + n].view_as(p))
            o += n
        return (None, vt.to(t.dtype).reshape(t.shape), vy) + tuple(grads)


class _LinComb(torch.autograd.Function):
    """out = base + sum_j coef[j] * src_j as ONE autograd node on the native kernels (csrc/lincomb.cu): the reference records one
    multiply and one add per term (misc.py:22-30); the forward rounds exactly like that op sequence, the backward is one pass for
    the sources' gradients (coef[j] * g) and one for the coefficients' (float64 dot products, fixed order)."""

    @staticmethod
    def forward(ctx, coef, base, *srcs):
        native = _solver.native
        lib = native.lib()
        srcs = tuple(v.contiguous() for v in srcs)
        ref = srcs[0]
        code = native.F32 if ref.dtype == torch.float32 else native.F64
        coef = coef.to(ref.dtype).contiguous()
        b = base.contiguous() if base is not None else None
        out = torch.empty_like(ref)
        arr = (native._vp * len(srcs))(*[v.data_ptr() for v in srcs])
        with native.device_guard(ref.device):
            native.check(lib.node_b200_lincomb(code, native.ptr(out), native.ptr(b), arr, native.ptr(coef), len(srcs), ref.numel(),
                                               native.stream_ptr()), 'lincomb')
        ctx.save_for_backward(coef, *srcs)
        ctx.code, ctx.has_base = code, base is not None
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        native = _solver.native
        lib = native.lib()
        coef, *srcs = ctx.saved_tensors
        g = g.contiguous()
        need = ctx.needs_input_grad
        n = len(srcs)
        with native.device_guard(g.device):
            gs = [torch.empty_like(g) if need[2 + j] else None for j in range(n)]
            if any(v is not None for v in gs):
                arr = (native._vp * n)(*[v.data_ptr() if v is not None else 0 for v in gs])
                native.check(lib.node_b200_lincomb_scale(ctx.code, arr, native.ptr(g), native.ptr(coef), n, g.numel(), native.stream_ptr()),
                             'lincomb_scale')
            gc = None
            if need[0]:
                gc = torch.empty(n, dtype=g.dtype, device=g.device)
                part = torch.empty(int(lib.node_b200_lincomb_scratch_doubles()), dtype=torch.float64, device=g.device)
                arr = (native._vp * n)(*[v.data_ptr() for v in srcs])
                native.check(lib.node_b200_lincomb_dots(ctx.code, arr, native.ptr(g), n, g.numel(), native.ptr(part), native.ptr(gc),
                                                        native.stream_ptr()), 'lincomb_dots')
        return (gc, g if (ctx.has_base and need[1]) else None) + tuple(gs)


_cvecs = {}


def _cvec(values, like):
    """The constants of a combination as a device vector in the state dtype (a Python double times a tensor rounds the double to the
    tensor's dtype first, so `h * _cvec(c)` equals the reference's `h * c` element by element)."""
    key = (tuple(values), like.dtype, str(like.device))
    v = _cvecs.get(key)
    if v is None:
        v = _cvecs[key] = torch.tensor([float(c) for c in values], dtype=like.dtype, device=like.device)
    return v


def _native_nodes(ys):
    import os
    return (os.environ.get('NODE_B200_UNROLLED_NODES', '1') != '0'
            and all(y.is_cuda and y.dtype in (torch.float32, torch.float64) and y.numel() > 0 for y in ys))


def _dynamics(func):
    """callable(t, tuple_state) -> tuple; the recognised ODE-Net dynamics go through the native kernels."""
    base = _solver._unwrap(func) if isinstance(func, nn.Module) else func
    params = _solver.recognise_odefunc(base) if isinstance(base, nn.Module) else None
    if params is not None:
        plist = tuple(base.parameters())

        def native_call(t, ys):
            y = ys[0]
            if len(ys) == 1 and _solver._fusable_state(params, (y,)) and _solver.native.lib().node_b200_vjp_workspace_bytes(
                    *[int(v) for v in y.shape]) > 0:
                if hasattr(base, 'nfe'):
                    base.nfe += 1                                  # model.py:340
                return (_NativeDynamics.apply(base, t, y, *plist),)
            return func(t, ys)
        return native_call
    return func


def _wsum(h, coeffs, ks):                     # misc.py:22-25: sum([(h*c)*k ...]): every product first, then the additions from
    return sum([(h * c) * k for c, k in zip(coeffs, ks)])      # int 0, left to right, zeros kept (autograd's summation order follows)


def _rms(x):                                  # misc.py:71-76
    return x.norm() / (x.numel() ** 0.5)


def _first_step(f, t0, ys, rtol, atol, f0):
    """misc.py:84-143 with order 4 (dopri5.py:80), in the state dtype."""
    t0 = t0.to(ys[0])
    scale = [atol + torch.abs(y) * rtol for y in ys]
    d0 = [_rms(y / s) for y, s in zip(ys, scale)]
    d1 = [_rms(v / s) for v, s in zip(f0, scale)]
    if max(d0).item() < 1e-5 or max(d1).item() < 1e-5:
        h0 = torch.tensor(1e-6).to(t0)
    else:
        h0 = 0.01 * max(a / b for a, b in zip(d0, d1))
    y1 = tuple(y + h0 * v for y, v in zip(ys, f0))
    f1 = f(t0 + h0, y1)
    d2 = [_rms((b - a) / s) / h0 for b, a, s in zip(f1, f0, scale)]
    if max(d1).item() <= 1e-15 and max(d2).item() <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6).to(h0), h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1 + d2)) ** (1.0 / 5.0)
    return torch.min(100 * h0, h1)


def _f64(x):                                  # misc.py:37-44: python number -> default dtype -> float64
    return torch.tensor(x).type(torch.float64)


def solve(func, y0, t, rtol, atol, options, stats=None):
    """y0: tuple of tensors, t: 1-D tensor (either direction); returns a tuple of [T, ...] tensors with the autograd graph of
    the reference's unrolled dopri5 (odeint.py:20-76, solvers.py:25-33, dopri5.py:60-122)."""
    f = _dynamics(func)
    ys = tuple(y0)
    if len(t) > 1 and bool((t[1:] < t[:-1]).all()):                  # misc.py:184-187
        t = -t
        fwd = f
        f = lambda tt, yy: tuple(-v for v in fwd(-tt, yy))
    n = len(ys)
    rtol = list(rtol) if _solver._is_iterable(rtol) else [rtol] * n  # dopri5.py:66-67
    atol = list(atol) if _solver._is_iterable(atol) else [atol] * n
    safety, ifactor, dfactor = (_f64(options.get(k, d)) for k, d in (('safety', 0.9), ('ifactor', 10.0), ('dfactor', 0.2)))
    max_steps = options.get('max_num_steps', 2 ** 31 - 1)
    dev = ys[0].device
    safety, ifactor, dfactor = safety.to(dev), ifactor.to(dev), dfactor.to(dev)
    t = t.to(dev, torch.float64)                                     # solvers.py:28
    k1 = f(t[0].type_as(ys[0]), ys)                                  # dopri5.py:78
    nfe = 1
    if options.get('first_step') is None:
        dt = _first_step(f, t[0], ys, rtol[0], atol[0], k1).to(t)    # dopri5.py:80
        nfe += 1
    else:
        dt = _f64(0.01).to(dev)                                      # dopri5.py:81-82
    t0 = t1 = t[0]
    coeffs = [ys] * 5
    outs = [ys]
    nat = _native_nodes(ys)            # CUDA states: every combination below is ONE native autograd node instead of 2 ATen ops per term
    n_acc = n_rej = 0
    for i in range(1, len(t)):
        steps = 0
        while bool(t[i] > t1):                                       # dopri5.py:88
            assert steps < max_steps, 'max_num_steps exceeded ({}>={})'.format(steps, max_steps)
            start = t1
            assert bool(start + dt > start), 'underflow in dt {}'.format(dt.item())
            for y in ys:
                assert bool(torch.isfinite(torch.abs(y)).all()), 'non-finite values in state `y`: {}'.format(y)
            h, s = dt.to(ys[0]), start.to(ys[0])                     # rk_common.py:45-46
            ks = [[v] for v in k1]
            yi = ys
            for a_i, b_i in zip(_ALPHA, _BETA):                      # rk_common.py:49-52
                ti = s + a_i * h                                     # node creation order as in the reference: it fixes
                if nat:
                    yi = tuple(_LinComb.apply(h * _cvec(b_i, y), y, *k) for y, k in zip(ys, ks))
                else:
                    yi = tuple(y + _wsum(h, b_i, k) for y, k in zip(ys, ks))   # the order in which autograd sums into h
                for k, v in zip(ks, f(ti, yi)):
                    k.append(v)
            nfe += 6
            y1, f1 = yi, tuple(k[-1] for k in ks)                    # FSAL (rk_common.py:54-58)
            ratios = []
            for k, a, b, rt, at in zip(ks, ys, y1, rtol, atol):      # misc.py:146-157
                err = _LinComb.apply(h * _cvec(_C_ERR, a), None, *k) if nat else _wsum(h, _C_ERR, k)
                q = err / (at + rt * torch.max(torch.abs(a), torch.abs(b)))
                ratios.append(torch.mean(q * q))
            accept = bool((torch.stack([r.detach() for r in ratios]) <= 1).all())     # dopri5.py:109
            if accept:
                t_next = start + dt                                                     # dopri5.py:112, before the fit
                hf = dt.type_as(ys[0])                                                  # dopri5.py:41: its own cast node
                if nat:
                    ymid = [_LinComb.apply(hf * _cvec(_C_MID, y), y, *k) for y, k in zip(ys, ks)]
                    fit = lambda ch, c0, m: tuple(_LinComb.apply(hf * _cvec(ch, v[0]) + _cvec(c0, v[0]), None, *v) for v in zip(k1, f1, ys, y1, m))
                    coeffs = [fit((-2, 2, 0, 0, 0), (0, 0, -8, -8, 16), ymid), fit((5, -3, 0, 0, 0), (0, 0, 18, 14, -32), ymid),
                              fit((-4, 1, 0, 0, 0), (0, 0, -11, -5, 16), ymid), tuple(hf * v for v in k1), ys]
                else:
                    ymid = [y + _wsum(hf, _C_MID, k) for y, k in zip(ys, ks)]           # dopri5.py:39-45, interp.py:5-35
                    fit = lambda cs, m: tuple(_lin(cs, v) for v in zip(k1, f1, ys, y1, m))
                    coeffs = [fit([-2 * hf, 2 * hf, -8, -8, 16], ymid), fit([5 * hf, -3 * hf, 18, 14, -32], ymid),
                              fit([-4 * hf, hf, -11, -5, 16], ymid), tuple(hf * v for v in k1), ys]
                t0, t1 = start, t_next
                ys, k1 = y1, f1
                n_acc += 1
            else:
                t0 = start                                           # dopri5.py:121 (the interpolant is kept)
                n_rej += 1
            r = max(ratios)                                          # misc.py:160-170
            if bool(r == 0):
                dt = dt * ifactor
            else:
                df = _f64(1).to(dev) if bool(r < 1) else dfactor
                factor = torch.max(1 / ifactor, torch.min(torch.sqrt(r).to(dt) ** torch.tensor(1 / 5).to(dt) / safety, 1 / df))
                dt = dt / factor
            steps += 1
        ref = coeffs[0][0]                                           # interp.py:38-65
        a0, a1, tt = t0.to(ref), t1.to(ref), t[i].to(ref)
        assert bool((a0 <= tt) & (tt <= a1)), 'invalid interpolation, fails `t0 <= t <= t1`: {}, {}, {}'.format(a0, tt, a1)
        x = (tt - a0) / (a1 - a0)
        pw = [torch.tensor(1).to(ref), x]
        for _ in range(2, 5):
            pw.append(pw[-1] * x)
        pw = pw[::-1]
        if nat:
            outs.append(tuple(_LinComb.apply(torch.stack(pw), None, *per) for per in zip(*coeffs)))
        else:
            outs.append(tuple(_lin(pw, per) for per in zip(*coeffs)))
    if stats is not None:
        stats.update(route='unrolled', nfe=nfe, n_accept=n_acc, n_reject=n_rej)
    return tuple(torch.stack(v) for v in zip(*outs))


def _lin(cs, xs):                              # misc.py:27-30 `_dot_product`
    return sum([c * x for c, x in zip(cs, xs)])


class LazyUnrolled(torch.autograd.Function):
    """`odeint` of the recognised ODE-Net dynamics with gradients enabled: forward = the fused solve (the route every no-grad call
    takes: one C call, no autograd graph, the reference's NFE), backward = the unrolled gradient of `solve` above, recorded and
    differentiated on demand (the replay re-integrates from y0 with the same arithmetic, so it takes the same step sequence; its
    evaluations are not counted in `func.nfe` a second time - the reference's backward does not evaluate `func` either). A model
    that is only ever called forward under enable_grad (evaluate.py:109-126, foolbox predictions) never pays for a graph."""

    @staticmethod
    def forward(ctx, module, call, t, rtol, atol, y0, *params):
        ctx.module, ctx.call, ctx.rtol, ctx.atol = module, call, rtol, atol
        with torch.no_grad():
            out = _solver._solve(call, (y0,), t, rtol, atol, {})[0]
        ctx.save_for_backward(t, y0)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        t, y0 = ctx.saved_tensors
        module = ctx.module
        params = tuple(module.parameters())
        nfe = getattr(module, 'nfe', None)
        stats = {}
        with torch.enable_grad(), torch.cuda.device(y0.device):
            y = y0.detach().requires_grad_(True)
            tt = t.detach().requires_grad_(True) if t.is_floating_point() else t
            out = solve(ctx.call, (y,), tt, ctx.rtol, ctx.atol, {}, stats=stats)[0]
            wanted = [y] + ([tt] if tt.requires_grad else []) + [p for p in params if p.requires_grad]
            grads = list(torch.autograd.grad(out, wanted, grad_out, allow_unused=True))
        if nfe is not None:
            module.nfe = nfe
        _solver.last_stats['grad_route'] = 'unrolled'
        _solver.last_stats['replay'] = stats
        gy = grads.pop(0)
        gt = grads.pop(0) if tt.requires_grad else None
        gp = [grads.pop(0) if p.requires_grad else None for p in params]
        return (None, None, gt, None, None, gy) + tuple(gp)
