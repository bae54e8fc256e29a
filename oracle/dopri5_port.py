"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement, in plain torch ops, of the adaptive Dormand-Prince 5(4) integrator
as implemented by the torchdiffeq snapshot pinned by fabiocarrara/neural-ode-features
(torchdiffeq @ a344d75).  Every routine cites the reference file:line it restates
(paths relative to /root/reference/torchdiffeq/torchdiffeq/_impl/).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (neural-ode-features_b200/) must not.

Parity status: PINNED.  tools/make_golden.py imports the unmodified reference in the
build container and checks this restatement against it bit-for-bit on CPU (state,
dt trace, accept/reject sequence, NFE) before writing tests/golden/*.npz;
tests/test_oracle.py re-checks the restatement against those committed vectors.

Arithmetic conventions preserved on purpose (they decide accept/reject):
  * time / dt / controller live in float64; they are cast to the state dtype
    before touching the state (rk_common.py:45-46, interp.py:54-56).
  * a weighted sum is evaluated left to right as ((h*c_0)*k_0 + (h*c_1)*k_1) + ...
    with h already in the state dtype (misc.py:22-25); zero coefficients are kept.
  * the error norm is a mean over the WHOLE tensor, batch included, one mean per
    member of a tuple state (misc.py:146-157); accept iff all means <= 1.
  * the solver never clips a step to land on an output time; outputs are 4th-order
    interpolants inside the last accepted step (dopri5.py:85-92).
"""
import torch

# Butcher tableau, dopri5.py:11-31 (values are the published Dormand-Prince /
# Shampine coefficients, written here as exact rationals evaluated in double).
ALPHA = (1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0)
BETA = (
    (1 / 5,),
    (3 / 40, 9 / 40),
    (44 / 45, -56 / 15, 32 / 9),
    (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
    (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656),
    (35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84),
)
C_ERR = (
    35 / 384 - 1951 / 21600,
    0,
    500 / 1113 - 22642 / 50085,
    125 / 192 - 451 / 720,
    -2187 / 6784 - -12231 / 42400,
    11 / 84 - 649 / 6300,
    -1.0 / 60.0,
)
# dense-output mid-point weights, dopri5.py:33-36
C_MID = (
    6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2,
    -2691868925 / 45128329728 / 2, 187940372067 / 1594534317056 / 2,
    -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2,
)


def weighted_sum(h, coeffs, ks):
    """misc.py:22-25: sum([(h*c)*k ...]) starting from int 0, left to right."""
    acc = 0
    for c, k in zip(coeffs, ks):
        acc = acc + (h * c) * k
    return acc


def rms(x):
    """misc.py:71-76 (tensor branch)."""
    return x.norm() / (x.numel() ** 0.5)


def is_finite(x):
    return bool(torch.isfinite(x).all())


def initial_step(func, t0, y0, order, rtol, atol, f0):
    """misc.py:84-143.  t0 is float64 0-d; everything below runs in the state dtype."""
    t0 = t0.to(y0[0])
    scale = [atol + torch.abs(y) * rtol for y in y0]
    d0 = [rms(y / s) for y, s in zip(y0, scale)]
    d1 = [rms(f / s) for f, s in zip(f0, scale)]
    if max(d0).item() < 1e-5 or max(d1).item() < 1e-5:
        h0 = torch.tensor(1e-6).to(t0)
    else:
        h0 = 0.01 * max(a / b for a, b in zip(d0, d1))
    y1 = [y + h0 * f for y, f in zip(y0, f0)]
    f1 = func(t0 + h0, y1)
    d2 = [rms((b - a) / s) / h0 for b, a, s in zip(f1, f0, scale)]
    if max(d1).item() <= 1e-15 and max(d2).item() <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6).to(h0), h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1 + d2)) ** (1.0 / float(order + 1))
    return torch.min(100 * h0, h1)


def error_ratios(err, rtol, atol, y0, y1):
    """misc.py:146-157."""
    out = []
    for e, a, b in zip(err, y0, y1):
        tol = atol + rtol * torch.max(torch.abs(a), torch.abs(b))
        q = e / tol
        out.append(torch.mean(q * q))
    return out


def _dflt(x):
    """misc.py:37-44 `_convert_to_tensor(x, dtype=float64)`: the Python number first becomes a tensor
    of the DEFAULT dtype (float32 unless the caller changed it) and is only then widened, so
    safety=0.9, dfactor=0.2 and the exponent 1/5 carry float32 rounding into the float64 controller."""
    return torch.tensor(x).type(torch.float64)


def next_step_size(dt, ratios, safety=0.9, ifactor=10.0, dfactor=0.2, order=5):
    """misc.py:160-170 with the solver's constants as built in dopri5.py:72-74."""
    r = max(ratios)
    if r == 0:
        return dt * _dflt(ifactor)
    df = _dflt(1) if r < 1 else _dflt(dfactor)
    e = torch.sqrt(r).to(dt)
    expo = torch.tensor(1 / order).to(dt)
    factor = torch.max(1 / _dflt(ifactor), torch.min(e ** expo / _dflt(safety), 1 / df))
    return dt / factor


def hermite_fit(y0, y1, ymid, f0, f1, dt):
    """interp.py:5-35; dt already in the state dtype."""
    def dot(cs, xs):
        acc = 0
        for c, x in zip(cs, xs):
            acc = acc + c * x
        return acc
    a = [dot([-2 * dt, 2 * dt, -8, -8, 16], v) for v in zip(f0, f1, y0, y1, ymid)]
    b = [dot([5 * dt, -3 * dt, 18, 14, -32], v) for v in zip(f0, f1, y0, y1, ymid)]
    c = [dot([-4 * dt, dt, -11, -5, 16], v) for v in zip(f0, f1, y0, y1, ymid)]
    d = [dt * f for f in f0]
    return [a, b, c, d, list(y0)]


def hermite_eval(coeffs, t0, t1, t):
    """interp.py:38-65: cast the three times to the state dtype first."""
    ref = coeffs[0][0]
    t0, t1, t = t0.to(ref), t1.to(ref), t.to(ref)
    assert (t0 <= t) & (t <= t1), 'invalid interpolation, fails `t0 <= t <= t1`: {}, {}, {}'.format(t0, t, t1)
    x = (t - t0) / (t1 - t0)
    one = torch.tensor(1).to(ref)
    pw = [one, x]
    for _ in range(2, len(coeffs)):
        pw.append(pw[-1] * x)
    pw = pw[::-1]
    out = []
    for per_tensor in zip(*coeffs):
        acc = 0
        for c, p in zip(per_tensor, pw):
            acc = acc + c * p
        out.append(acc)
    return out


class Trace(object):
    """What the parity tests compare: one record per attempted step."""

    def __init__(self):
        self.steps = []      # (t_start, dt, accepted, [ratio per tensor])
        self.nfe = 0
        self.dt0 = None

    @property
    def n_accept(self):
        return sum(1 for s in self.steps if s[2])

    @property
    def n_reject(self):
        return sum(1 for s in self.steps if not s[2])


def dopri5_solve(func, y0, t, rtol, atol, trace=None, max_num_steps=2 ** 31 - 1,
                 norm_reduce=None):
    """odeint.py:20-76 + solvers.py:25-33 + dopri5.py:60-122 for method='dopri5'.

    func(t, [tensors]) -> [tensors];  y0: tensor or tuple/list of tensors;  t: 1-D.
    norm_reduce (oracle extension used by the world_size>1 host-logic tests): a
    callable(sum_of_squares_tensor, numel) -> (global_sum, global_numel) that stands
    in for the all-reduce of SURVEY 8(e); None means single shard.
    """
    single = torch.is_tensor(y0)
    ys = [y0] if single else list(y0)
    user = func
    if single:
        func = lambda tt, yy: [user(tt, yy[0])]
    else:
        func = lambda tt, yy: list(user(tt, tuple(yy)))
    if bool((t[1:] < t[:-1]).all()) and len(t) > 1:      # misc.py:184-187
        t = -t
        fwd = func
        func = lambda tt, yy: [-v for v in fwd(-tt, yy)]
    for y in ys:
        if not torch.is_floating_point(y):
            raise TypeError('`y0` must be a floating point Tensor but is a {}'.format(y.type()))
    if not torch.is_floating_point(t):
        raise TypeError('`t` must be a floating point Tensor but is a {}'.format(t.type()))
    assert bool((t[1:] > t[:-1]).all()), 't must be strictly increasing or decrasing'
    tr = trace if trace is not None else Trace()

    def counted(tt, yy):
        tr.nfe += 1
        return func(tt, yy)

    t = t.to(ys[0].device, torch.float64)
    f = counted(t[0].type_as(ys[0]), ys)                 # dopri5.py:78
    if norm_reduce is None:
        dt = initial_step(counted, t[0], ys, 4, rtol, atol, f).to(t)
    else:
        dt = _initial_step_sharded(counted, t[0], ys, 4, rtol, atol, f, norm_reduce).to(t)
    tr.dt0 = float(dt)
    t0 = t1 = t[0]
    coeffs = [list(ys)] * 5
    outputs = [list(ys)]
    for i in range(1, len(t)):
        nsteps = 0
        while t[i] > t1:                                  # dopri5.py:88
            assert nsteps < max_num_steps, 'max_num_steps exceeded ({}>={})'.format(nsteps, max_num_steps)
            start = t1
            assert start + dt > start, 'underflow in dt {}'.format(dt.item())
            for y in ys:
                assert is_finite(torch.abs(y)), 'non-finite values in state `y`: {}'.format(y)
            h = dt.to(ys[0])                              # rk_common.py:45-46
            s = start.to(ys[0])
            ks = [[v] for v in f]
            yi = ys
            for a_i, b_i in zip(ALPHA, BETA):             # rk_common.py:49-52
                ti = s + a_i * h
                yi = [y + weighted_sum(h, b_i, k) for y, k in zip(ys, ks)]
                for k, v in zip(ks, counted(ti, yi)):
                    k.append(v)
            y1 = yi                                       # FSAL, rk_common.py:54-58
            f1 = [k[-1] for k in ks]
            err = [weighted_sum(h, C_ERR, k) for k in ks]
            if norm_reduce is None:
                ratios = error_ratios(err, rtol, atol, ys, y1)
            else:
                ratios = _error_ratios_sharded(err, rtol, atol, ys, y1, norm_reduce)
            ok = bool((torch.tensor(ratios) <= 1).all())  # dopri5.py:109
            tr.steps.append((float(start), float(dt), ok, [float(r) for r in ratios]))
            if ok:
                ymid = [y + weighted_sum(h, C_MID, k) for y, k in zip(ys, ks)]
                coeffs = hermite_fit(ys, y1, ymid, [k[0] for k in ks], f1, h)
                t0, t1 = start, start + dt
                ys, f = y1, f1
            else:
                t0 = start
            dt = next_step_size(dt, ratios)
            nsteps += 1
        outputs.append(hermite_eval(coeffs, t0, t1, t[i]))
    stacked = [torch.stack(v) for v in zip(*outputs)]
    return stacked[0] if single else tuple(stacked)


# ---- sharded-norm variants (host-logic oracle for SURVEY 8(e)) ---------------------------

def _rms_sharded(x, norm_reduce):
    ssq, n = norm_reduce((x.double() ** 2).sum(), x.numel())
    return torch.sqrt(ssq / n).to(x.dtype)


def _initial_step_sharded(func, t0, y0, order, rtol, atol, f0, norm_reduce):
    t0 = t0.to(y0[0])
    scale = [atol + torch.abs(y) * rtol for y in y0]
    d0 = [_rms_sharded(y / s, norm_reduce) for y, s in zip(y0, scale)]
    d1 = [_rms_sharded(f / s, norm_reduce) for f, s in zip(f0, scale)]
    if max(d0).item() < 1e-5 or max(d1).item() < 1e-5:
        h0 = torch.tensor(1e-6).to(t0)
    else:
        h0 = 0.01 * max(a / b for a, b in zip(d0, d1))
    y1 = [y + h0 * f for y, f in zip(y0, f0)]
    f1 = func(t0 + h0, y1)
    d2 = [_rms_sharded((b - a) / s, norm_reduce) / h0 for b, a, s in zip(f1, f0, scale)]
    if max(d1).item() <= 1e-15 and max(d2).item() <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6).to(h0), h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1 + d2)) ** (1.0 / float(order + 1))
    return torch.min(100 * h0, h1)


def _error_ratios_sharded(err, rtol, atol, y0, y1, norm_reduce):
    out = []
    for e, a, b in zip(err, y0, y1):
        tol = atol + rtol * torch.max(torch.abs(a), torch.abs(b))
        q = e / tol
        ssq, n = norm_reduce((q.double() * q.double()).sum(), q.numel())
        out.append((ssq / n).to(e.dtype))
    return out


# ---- adjoint (adjoint.py:7-133) -----------------------------------------------------------

def flat_params(params):
    v = [p.contiguous().view(-1) for p in params]
    return torch.cat(v) if v else torch.tensor([])


def adjoint_backward(func, params, t, ys, grad_ys, rtol, atol, vjp=None, trace=None):
    """adjoint.py:23-102 for a single-tensor state.

    func(t, y) -> f; params: tuple of tensors func depends on;
    ys, grad_ys: [T, *shape].  vjp(t, y, a) -> (f, vjp_y, vjp_t, flat vjp_params) overrides
    the autograd-based VJP (used to pin the hand-derived ODEfunc backward).
    Returns (grad_y0, grad_t[T], grad_flat_params).
    """
    params = tuple(params)

    def autograd_vjp(tt, y, a):
        with torch.enable_grad():
            tt = tt.detach().requires_grad_(True)
            y = y.detach().requires_grad_(True)
            fe = func(tt, y)
            g = torch.autograd.grad(fe, (tt, y) + params, a, allow_unused=True)
        vt = torch.zeros_like(tt) if g[0] is None else g[0]
        vy = torch.zeros_like(y) if g[1] is None else g[1]
        vp = [torch.zeros_like(p).view(-1) if q is None else q.contiguous().view(-1)
              for q, p in zip(g[2:], params)]
        vp = torch.cat(vp) if vp else torch.tensor(0.).to(vy)
        return fe.detach(), vy, vt, vp

    vjp = vjp or autograd_vjp

    def aug(tt, state):                                   # adjoint.py:32-55
        y, a = state[0], state[1]
        fe, vy, vt, vp = vjp(tt.to(y.device), y, -a)
        return (fe, vy, vt, vp)

    T = ys.shape[0]
    with torch.no_grad():
        adj_y = grad_ys[-1]
        adj_p = torch.zeros_like(flat_params(params))
        adj_t = torch.tensor(0.).to(t)
        tv = []
        for i in range(T - 1, 0, -1):
            fi = func(t[i], ys[i])
            d = torch.dot(fi.reshape(-1), grad_ys[i].reshape(-1)).view(1)
            adj_t = adj_t - d
            tv.append(d)
            if len(adj_p) == 0:
                adj_p = torch.tensor(0.).to(adj_y)
            sol = dopri5_solve(aug, (ys[i], adj_y, adj_t, adj_p), torch.tensor([t[i], t[i - 1]]),
                               rtol, atol, trace=trace)
            adj_y, adj_t, adj_p = sol[1][1], sol[2][1], sol[3][1]
            adj_y = adj_y + grad_ys[i - 1]
        tv.append(adj_t)
        return adj_y, torch.cat(tv[::-1]), adj_p
