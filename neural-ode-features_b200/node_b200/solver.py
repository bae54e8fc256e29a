"""`odeint` / `odeint_adjoint` with the pinned signatures of the reference's torchdiffeq snapshot
(/root/reference/torchdiffeq/torchdiffeq/_impl/odeint.py:20, adjoint.py:105), served by the sm_100a
kernels in csrc/ through the C ABI (include/node_b200.h).

Two routes, both device-only (no CPU path, no eager fallback):
  * fused route - `func` is the ODE-Net dynamics module (model.py:326-348, 64 filters, GroupNorm):
    the whole solve is enqueued by one C call; dynamics, RK stages, error norm, controller and
    dense output all run on the device, the host reads the controller block back once per solve.
  * generic route - any callable, float32 or float64, tensor or tuple state, either time
    direction: the callable is evaluated by the caller's own code (it is host Python by
    definition); stage combination, error norm, controller, initial step and dense output are the
    CUDA kernels K2-K6, with one host read of the controller block per attempted step instead of
    the reference's >= 9 syncs.
Around them (round 2): the wide route (128 / 192 / 256 filters: wide.py), the adjoint's backward interval
as one C call with a device-side loop (node_b200_adjoint_solve), gradients through the non-adjoint
`odeint` (unrolled.py: the reference's unrolled gradient, fused forward + lazily recorded replay for the
recognised dynamics), N independent per-sample solves (each.py).
"""
import os
import warnings
import weakref

import numpy as np
import torch
import torch.nn as nn

from . import native
from . import distributed as dist_state

_KNOWN_METHODS = ('explicit_adams', 'fixed_adams', 'adams', 'tsit5', 'dopri5', 'euler', 'midpoint', 'rk4')
_DOPRI5_OPTIONS = ('first_step', 'safety', 'ifactor', 'dfactor', 'max_num_steps')
CONV_MODES = {'f16x3': 0, 'f16': 1, 'simt': 2, 'tf32x3': 3, 'tf32': 4}

last_stats = {}          # filled after every solve: nfe, n_accept, n_reject, trace, route, launches
# bench.py sets this to a list: the fused route then records a (start, end) CUDA-event pair around
# every launch of the 6-stage step kernel on the launching stream (roofline timing, live)
PROFILE_STEP_EVENTS = None
_t_cache = {}
_ws_cache = {}
class _nvtx(object):
    """NVTX range around the solver's entry points (SURVEY 5 tracing) when NODE_B200_NVTX=1: nsys / ncu --nvtx timelines then
    show `node_b200.solve`, `node_b200.adjoint_backward`, `node_b200.unrolled` around the kernels they enqueue."""

    def __init__(self, name):
        self.name = name
        self.on = os.environ.get('NODE_B200_NVTX', '0') == '1' and torch.cuda.is_available()

    def __enter__(self):
        if self.on:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *a):
        if self.on:
            torch.cuda.nvtx.range_pop()
        return False


class _StepGuess(object):
    """Attempted-step counts of the last solve per dynamics object (sizes the enqueue-ahead / captured graphs). Keyed by the
    object itself through a weak reference, so an entry dies with its module instead of being inherited by whatever object
    is allocated at the same address next."""

    def __init__(self):
        self._d = weakref.WeakKeyDictionary()
        self._plain = {}

    def _slot(self, key):
        tag, obj = key if isinstance(key, tuple) else (None, key)
        try:
            per = self._d.setdefault(obj, {})
        except TypeError:                      # not weak-referenceable (a builtin callable): fall back to its id
            per = self._plain.setdefault(id(obj), {})
        return per, tag

    def get(self, key, default=None):
        per, tag = self._slot(key)
        return per.get(tag, default)

    def __setitem__(self, key, value):
        per, tag = self._slot(key)
        per[tag] = value

    def clear(self):
        self._d = weakref.WeakKeyDictionary()
        self._plain = {}


_step_guess = _StepGuess()
_slow_shape_warned = set()
_warned_grad_bridge = False


def _conv_mode():
    name = os.environ.get('NODE_B200_CONV', 'f16x3')
    if name not in CONV_MODES:
        raise ValueError('NODE_B200_CONV must be one of %s' % sorted(CONV_MODES))
    return CONV_MODES[name]


def _host_times(t):
    """float64 host copy of `t` (one sync the first time a given tensor is seen). Freshness is judged by (identity, _version):
    a tensor that requires grad is never cached (torch.autograd.gradcheck perturbs it through `.data`, which bumps no version)."""
    if t.requires_grad:
        return t.detach().to('cpu', torch.float64).numpy().copy()
    key = id(t)
    hit = _t_cache.get(key)
    if hit is not None and hit[0]() is t and hit[1] == t._version:
        return hit[2]
    host = t.detach().to('cpu', torch.float64).numpy().copy()
    if len(_t_cache) > 64:
        _t_cache.clear()
    try:
        _t_cache[key] = (weakref.ref(t), t._version, host)
    except TypeError:
        pass
    return host


def _check_inputs(func, y0, t):
    """misc.py:173-195 minus the reversal (handled per route)."""
    tensor_input = torch.is_tensor(y0)
    if tensor_input:
        y0 = (y0,)
    assert isinstance(y0, tuple), 'y0 must be either a torch.Tensor or a tuple'
    for y in y0:
        assert torch.is_tensor(y), 'each element must be a torch.Tensor but received {}'.format(type(y))
    for y in y0:
        if not torch.is_floating_point(y):
            raise TypeError('`y0` must be a floating point Tensor but is a {}'.format(y.type()))
    if not torch.is_floating_point(t):
        raise TypeError('`t` must be a floating point Tensor but is a {}'.format(t.type()))
    return tensor_input, y0


def _needs_grad(func, y0, t):
    if not torch.is_grad_enabled():
        return False
    if any(y.requires_grad for y in y0) or t.requires_grad:
        return True
    if isinstance(func, nn.Module):
        return any(p.requires_grad for p in func.parameters())
    return any(p.requires_grad for m in _closure_modules(func) for p in m.parameters())


def _closure_modules(fn):
    """nn.Modules a plain callable can reach: its closure cells, its default arguments, the object a bound method belongs to,
    and the same one level down for functions found there (`lambda t, y: (f(t, y[0]), f(t, y[1]))` is the api_tests.py case)."""
    found, seen = [], set()

    def visit(obj, depth):
        if id(obj) in seen:
            return
        seen.add(id(obj))
        if isinstance(obj, nn.Module):
            found.append(obj)
            return
        if depth > 2:
            return
        owner = getattr(obj, '__self__', None)
        if owner is not None:
            visit(owner, depth + 1)
        for cell in getattr(obj, '__closure__', None) or ():
            try:
                visit(cell.cell_contents, depth + 1)
            except ValueError:                       # empty cell
                pass
        for d in getattr(obj, '__defaults__', None) or ():
            visit(d, depth + 1)
        if isinstance(obj, (list, tuple)):
            for o in obj:
                visit(o, depth + 1)

    visit(fn, 0)
    return found


class _CallableModule(nn.Module):
    """A plain callable as an nn.Module, so that gradients requested through `odeint` can be served by the adjoint ODE
    (which needs parameters to differentiate with respect to): the modules the callable closes over become sub-modules."""

    def __init__(self, fn):
        super().__init__()
        self._fn = fn
        self.reached = nn.ModuleList(_closure_modules(fn))

    def forward(self, t, y):
        return self._fn(t, y)


class _TensorFunc(nn.Module):
    """Tuple adapter (adjoint.py:113-126) that keeps the wrapped dynamics module visible to the recogniser."""

    def __init__(self, base):
        super().__init__()
        self.base = base

    def forward(self, t, y):
        return (self.base(t, y[0]),)


def _unwrap(func):
    return func.base if isinstance(func, _TensorFunc) else func


def recognise_odefunc(func):
    """Duck-typed recogniser of the ODE-Net dynamics (model.py:326-348). Returns its parameters in
    kernel order (conv1 W, b, conv2 W, b, norm1 w, b, norm2 w, b, norm3 w, b) or None."""
    func = _unwrap(func)
    if os.environ.get('NODE_B200_FUSED', '1') == '0' or not isinstance(func, nn.Module):
        return None
    if type(func).__name__ != 'ODEfunc' and not getattr(func, '_node_b200_fusable', False):
        return None
    try:
        convs = [func.conv1._layer, func.conv2._layer]
        norms = [func.norm1, func.norm2, func.norm3]
    except AttributeError:
        return None
    if not isinstance(getattr(func, 'relu', None), nn.ReLU):
        return None
    C = convs[0].out_channels
    for c in convs:
        if type(c) is not nn.Conv2d or c.in_channels != C + 1 or c.out_channels != C or c.kernel_size != (3, 3) \
                or c.stride != (1, 1) or c.padding != (1, 1) or c.dilation != (1, 1) or c.groups != 1 \
                or c.bias is None or c.padding_mode != 'zeros':
            return None
    for n in norms:
        if type(n) is not nn.GroupNorm or n.num_groups != min(32, C) or n.num_channels != C or not n.affine \
                or abs(n.eps - 1e-5) > 1e-12:
            return None
    if C != 64:
        return None
    params = [convs[0].weight, convs[0].bias, convs[1].weight, convs[1].bias,
              norms[0].weight, norms[0].bias, norms[1].weight, norms[1].bias, norms[2].weight, norms[2].bias]
    for p in params:
        if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
            return None
    return params


def _fusable_state(params, y0):
    if len(y0) != 1:
        return False
    y = y0[0]
    if y.dtype != torch.float32 or not y.is_cuda or y.dim() != 4 or y.shape[1] != 64:
        return False
    if y.device != params[0].device:
        return False
    return native.lib().node_b200_fused_workspace_bytes(int(y.shape[0]), 64, int(y.shape[2]), int(y.shape[3])) > 0


# ---- fused route ---------------------------------------------------------------------------------

class FusedWorkspace(object):
    def __init__(self, device, N, C, H, W):
        nbytes = native.lib().node_b200_fused_workspace_bytes(N, C, H, W)
        if nbytes <= 0:
            raise ValueError('shape [%d,%d,%d,%d] is not supported by the fused ODEfunc kernels' % (N, C, H, W))
        self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        self.shape = (N, C, H, W)
        self.param_key = None
        self.param_key_fresh = False
        L = native.layout()
        ctl_ptr = native.lib().node_b200_fused_ctl(native.ptr(self.buf))
        off = ctl_ptr - self.buf.data_ptr()
        self.ctl = self.buf[off:off + L['sizeof']]
        sums_ptr = native.lib().node_b200_fused_sums(native.ptr(self.buf))
        off = sums_ptr - self.buf.data_ptr()
        self.sums = self.buf[off:off + 8 * 2 * L['max_seg']].view(torch.float64)

    def prepare(self, params):
        # (address, version) alone can repeat when a model is freed and another one is built in the same process
        # (evaluate.py walks over runs): the weak references pin the identity of the live nn.Parameters as well.
        key = tuple((p.data_ptr(), p._version) for p in params)
        refs = getattr(self, 'param_refs', None)
        same = refs is not None and len(refs) == len(params) and all(r() is p for r, p in zip(refs, params))
        if key == self.param_key and same:
            return
        N, C, H, W = self.shape
        err = native.lib().node_b200_fused_prepare(native.ptr(self.buf), C, H, W, *[native.ptr(p) for p in params],
                                                  1e-5, native.stream_ptr())
        native.check(err, 'fused_prepare')
        self.param_key = key
        self.param_refs = [weakref.ref(p) for p in params]
        self.param_key_fresh = True


def invalidate_caches():
    """Forget everything that is keyed by tensor identity and `_version`: host copies of time tensors, prepared weight tiles
    (solver and caller kernels), captured CUDA graphs. In-place writes through `.data` (EMA / SWA swaps, `p.data.copy_()`, weight
    clipping) bump no version: call this after them."""
    _t_cache.clear()
    _graph_cache.clear()
    _graph_failed.clear()
    for ws in _ws_cache.values():
        ws.param_key = None
        ws.param_refs = None
    from . import caller_ops, wide
    caller_ops._resconv_ws.clear()
    caller_ops._convs2_ws.clear()
    for dyn in list(wide.WideDynamics._cache.values()):          # prepared weight images of the wide dynamics (forward and adjoint)
        for key in ('_key', '_key8', '_key8d', '_key_vjp', '_tmap8_key'):
            if hasattr(dyn, key):
                delattr(dyn, key)
    if _adjoint_bufs:
        _adjoint_bufs.clear()
        native.lib().node_b200_adjoint_solve_reset()             # loop graphs point at the buffers just released


def fused_workspace(device, N, C, H, W):
    key = (str(device), N, C, H, W)
    ws = _ws_cache.get(key)
    if ws is None:
        if len(_ws_cache) > 8:
            _ws_cache.clear()
            _graph_cache.clear()          # captured graphs point into the workspaces that are being released
            _graph_failed.clear()
        ws = _ws_cache[key] = FusedWorkspace(device, N, C, H, W)
    return ws


@native.on_device_of(2)
def odefunc_forward(func, t, y, tsign=1.0, conv_mode=None):
    """One evaluation of the recognised dynamics by the fused kernel (used by tests / the adjoint)."""
    params = recognise_odefunc(func)
    if params is None or not _fusable_state(params, (y,)):
        raise ValueError('func / y are not served by the fused ODEfunc kernels')
    y = y.contiguous()
    N, C, H, W = y.shape
    ws = fused_workspace(y.device, N, C, H, W)
    ws.prepare(params)
    k = torch.empty_like(y)
    err = native.lib().node_b200_odefunc_forward(native.ptr(ws.buf), native.ptr(y), float(t), float(tsign), native.ptr(k),
                                                N, C, H, W, _conv_mode() if conv_mode is None else conv_mode,
                                                native.stream_ptr())
    native.check(err, 'odefunc_forward')
    return k


_vjp_ws_cache = {}
_adjoint_bufs = {}              # (device, row length, T) -> state rows / controller / sums of node_b200_adjoint_solve
_adjoint_solve_failed = False
N_PARAMS_64 = 2 * (64 * 65 * 9 + 64) + 6 * 64


def _vjp_workspace(device, N, C, H, W, slot=0):
    key = (str(device), N, C, H, W, slot)
    buf = _vjp_ws_cache.get(key)
    if buf is None:
        nbytes = native.lib().node_b200_vjp_workspace_bytes(N, C, H, W)
        if nbytes <= 0:
            raise ValueError('shape [%d,%d,%d,%d] is not supported by the fused VJP kernels' % (N, C, H, W))
        if len(_vjp_ws_cache) > 8:
            _vjp_ws_cache.clear()
        buf = _vjp_ws_cache[key] = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    return buf


@native.on_device_of(2)
def odefunc_vjp(func, t, y, adj_y, tsign=1.0, out=None):
    """One evaluation of the adjoint's augmented dynamics (adjoint.py:32-55) by the native kernels:
    returns (f, vjp_y, vjp_t, vjp_params) with cotangent -adj_y, all multiplied by tsign and evaluated at
    tsign*t. `t` is a python float or a 0-d float32 CUDA tensor; `out` optionally names the four destination tensors."""
    params = recognise_odefunc(func)
    if params is None or not _fusable_state(params, (y,)):
        raise ValueError('func / y are not served by the fused ODEfunc kernels')
    N, C, H, W = y.shape
    ws = fused_workspace(y.device, N, C, H, W)
    ws.prepare(params)
    vws = _vjp_workspace(y.device, N, C, H, W)
    if not torch.is_tensor(t):
        t = torch.tensor(float(t), dtype=torch.float32, device=y.device)
    assert t.dtype == torch.float32 and t.is_cuda
    if out is None:
        out = (torch.empty_like(y), torch.empty_like(y), torch.empty((), dtype=y.dtype, device=y.device),
               torch.empty(N_PARAMS_64, dtype=y.dtype, device=y.device))
    f, vy, vt, vp = out
    for o in (y, adj_y, f, vy, vp):
        assert o.is_contiguous()
    err = native.lib().node_b200_odefunc_vjp(native.ptr(ws.buf), native.ptr(vws), native.ptr(y), native.ptr(adj_y),
                                            native.ptr(t), float(tsign), native.ptr(f), native.ptr(vy), native.ptr(vt),
                                            native.ptr(vp), N, C, H, W, native.stream_ptr())
    native.check(err, 'odefunc_vjp')
    return out


class _FusedAugmented(object):
    """The adjoint's augmented dynamics (adjoint.py:32-55) served by the native VJP kernels: the generic route
    calls eval_into() with views of its own stage buffers, so nothing is copied and no autograd graph exists."""

    # Members 2.. of the augmented state (adj_t, adj_params) are SUMS over the batch: when the batch is sharded every
    # rank holds a partial; the generic route all-reduces them after each evaluation so that all ranks integrate the
    # identical global values, and counts their error norms once (SURVEY 8e "adjoint collective").
    replicated_from = 2

    def __init__(self, func):
        self.func = func
        self.target = _unwrap(func)

    def eval_into(self, t_dev, src, dst, tsign):
        y, adj_y = src[0], src[1]
        odefunc_vjp(self.func, t_dev, y, adj_y, tsign=tsign, out=(dst[0], dst[1], dst[2], dst[3]))
        if hasattr(self.target, 'nfe'):
            self.target.nfe += 1                  # model.py:340 counts every evaluation


def _solve_fused(func, params, y0, t_host, tsign, rtol, atol):
    y = y0.detach().contiguous()
    N, C, H, W = y.shape
    T = len(t_host)
    ws = fused_workspace(y.device, N, C, H, W)
    ws.prepare(params)
    out = torch.empty((T,) + tuple(y.shape), dtype=y.dtype, device=y.device)
    lib = native.lib()
    if (int(H), int(W)) not in _slow_shape_warned and lib.node_b200_vjp_workspace_bytes(int(N), int(C), int(H), int(W)) <= 0:
        _slow_shape_warned.add((int(H), int(W)))
        warnings.warn('node_b200: %dx%d feature maps are served by the first-generation fused kernel (about 3x slower per step than '
                      'the tcgen05 step engine, which is tiled for 6x6, 7x7, 8x8, 14x14 and 16x16 maps)' % (H, W))
    conv_mode = _conv_mode()
    group = dist_state.group()
    th = native.host_f64(t_host)
    E = y.numel()
    common = (th, T, float(rtol), float(atol), N, C, H, W)
    launches = 7 + (1 if ws.param_key_fresh else 0)
    ws.param_key_fresh = False
    events = PROFILE_STEP_EVENTS
    peer = group is not None and dist_state.peer_active()          # sums all-reduced inside the fold kernel (peer memory)
    if (group is None or peer) and events is None:
        guess = _step_guess.get(_unwrap(func))
        graph = None
        if guess is not None and _graph_wanted(N, T) and not peer:
            # Launch-bound sizes: the whole enqueue sequence (f0, probe, `guess` attempted steps with their controllers
            # and dense output) is captured ONCE per (shape, times, tolerances, step count) and replayed.
            graph = _fused_graph(ws, y, th, t_host, common, E, conv_mode, int(tsign), guess)
        first = 1
        if graph is not None:
            graph.y.copy_(y)
            graph.graph.replay()
            src_y, dst_out = graph.y, graph.out
            launches += 4 * guess
            view = native.CtlView(ws.ctl)
            first = 0 if view.i32('done') else -1
        else:
            src_y, dst_out = y, out
            guess = 8 if guess is None else guess
        while first != 0:
            if first == -1:
                first, guess = 0, 4
            err = lib.node_b200_fused_solve(native.ptr(ws.buf), native.ptr(src_y), *common, E, native.ptr(dst_out), conv_mode,
                                            int(tsign), first, guess, native.stream_ptr())
            native.check(err, 'fused_solve')
            launches += 4 * guess
            view = native.CtlView(ws.ctl)
            if view.i32('done'):
                break
            first = -1
        if graph is not None:
            out.copy_(graph.out)
    else:
        sharded = group is not None
        # peer mode: the global element count is reduced on the device with the first norms (no collective, no read-back here)
        E_glob = dist_state.global_numel(E, y.device) if (sharded and not peer) else E

        def phase(ph):
            native.check(lib.node_b200_fused_phase(native.ptr(ws.buf), ph, native.ptr(y), *common, E_glob, native.ptr(out),
                                                   conv_mode, int(tsign), native.stream_ptr()), 'fused_phase %d' % ph)

        def reduce():
            if sharded and not peer:                                # peer mode: phases 0, 1, 3 end with the reduced sums
                dist_state.all_reduce_sum(ws.sums)

        phase(0); reduce()
        phase(1); reduce()
        phase(2)
        guess = _step_guess.get(_unwrap(func), 8)
        mine = []
        while True:
            for _ in range(guess):
                if events is not None:
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); phase(3); b.record()
                    mine.append((a, b))
                else:
                    phase(3)
                reduce()
                phase(4)
            launches += 4 * guess
            view = native.CtlView(ws.ctl)
            if view.i32('done'):       # identical on every rank: the controller consumed identical sums
                break
            guess = 4
        if events is not None:         # keep only launches that did work (steps after `done` are no-ops)
            events.extend(mine[:view.i32('n_attempt')])
    _step_guess[_unwrap(func)] = view.i32('n_attempt') + 1
    nfe = view.i32('nfe')
    target = _unwrap(func)
    if hasattr(target, 'nfe'):
        target.nfe += nfe              # model.py:340 counts one per evaluation; callers read it (train.py:49)
    last_stats.clear()
    last_stats.update(route='fused', nfe=nfe, n_accept=view.i32('n_accept'), n_reject=view.i32('n_reject'),
                      status=view.i32('status'), trace=view.trace(), launches=launches)
    native.raise_for_status(view.i32('status'))
    return out


# ---- CUDA-graph replay of the fused solve -------------------------------------------------------

_graph_cache = {}
_graph_failed = set()


def _graph_wanted(N, T):
    mode = os.environ.get('NODE_B200_GRAPH', 'auto')
    if mode == '0' or T > 32:
        return False
    return mode == '1' or N <= 512          # auto: sizes at which a solve is bound by its ~35 launches


class _FusedGraph(object):
    def __init__(self, y, T):
        self.y = torch.empty_like(y)
        self.out = torch.empty((T,) + tuple(y.shape), dtype=y.dtype, device=y.device)
        self.graph = torch.cuda.CUDAGraph()


def _fused_graph(ws, y, th, t_host, common, E, conv_mode, tsign, guess):
    key = (id(ws), ws.shape, str(y.device), tuple(float(v) for v in t_host), common[2], common[3], conv_mode, tsign, guess,
           os.environ.get('NODE_B200_STEP8'))          # the captured launches are engine-specific
    g = _graph_cache.get(key)
    if g is not None or key in _graph_failed:
        return g
    if len(_graph_cache) >= 16:
        _graph_cache.clear()
    g = _FusedGraph(y, len(t_host))
    g.ws = ws                                 # the graph's kernels point into this workspace: keep it alive with the graph
    try:
        with torch.cuda.graph(g.graph):
            err = native.lib().node_b200_fused_solve(native.ptr(ws.buf), native.ptr(g.y), *common, E, native.ptr(g.out),
                                                    conv_mode, tsign, 1, guess, native.stream_ptr())
        native.check(err, 'fused_solve (graph capture)')
    except Exception as exc:                      # same kernels without the graph; say so once per configuration
        _graph_failed.add(key)
        warnings.warn('node_b200: CUDA-graph capture of the fused solve failed (%s); launching directly' % exc)
        return None
    _graph_cache[key] = g
    return g


# ---- generic route -------------------------------------------------------------------------------

class _GenericSolve(object):
    NBUF = 11  # Y0 Y1 F0 F1 K2..K6 YMID YI

    def __init__(self, func, y0, t_host, rtol, atol, opts, tsign=1, rep_from=None):
        self.func = func
        self.tsign = tsign
        self.rep_from = len(y0) if rep_from is None else rep_from   # first member that is a batch sum (identical on all ranks)
        ref = y0[0]
        if ref.dtype not in (torch.float32, torch.float64):
            raise TypeError('node_b200 serves float32 and float64 states, got {}'.format(ref.dtype))
        for y in y0:
            if y.dtype != ref.dtype or y.device != ref.device:
                raise TypeError('all members of a tuple state must share dtype and device')
        self.dtype, self.device = ref.dtype, ref.device
        self.code = native.F32 if ref.dtype == torch.float32 else native.F64
        self.shapes = [tuple(y.shape) for y in y0]
        self.lens = [int(y.numel()) for y in y0]
        if len(y0) > native.layout()['max_seg']:
            raise ValueError('tuple states with more than %d members are not supported' % native.layout()['max_seg'])
        offs, o = [], 0
        for n in self.lens:
            offs.append(o)
            o += (n + 7) // 8 * 8
        self.offs, self.L = offs, max(o, 8)
        self.T = len(t_host)
        self.t_host = native.host_f64(t_host)
        L = native.layout()
        # The adjoint of the fused ODE-Net dynamics on one GPU runs as ONE C call per interval (node_b200_adjoint_solve): its
        # loop graph points at these buffers, so they are kept (and reused) per state shape instead of allocated per solve.
        self.one_call = (isinstance(func, _FusedAugmented) and dist_state.group() is None and self.code == native.F32
                         and len(y0) == 4 and len(self.shapes[0]) == 4 and os.environ.get('NODE_B200_ADJOINT_SOLVE', '1') != '0'
                         and not _adjoint_solve_failed)
        kept = _adjoint_bufs.get((str(self.device), self.L, self.T)) if self.one_call else None
        if kept is not None:
            self.bufs, self.ctl, self.partials, self.sums, self.flag, self.t_dev, self.out = kept
            self.t_dev.copy_(torch.tensor([float(v) for v in t_host], dtype=torch.float64))
        else:
            self.bufs = torch.zeros(self.NBUF, self.L, dtype=self.dtype, device=self.device)
            self.ctl = torch.zeros(L['sizeof'], dtype=torch.uint8, device=self.device)
            self.partials = torch.zeros(2 * L['max_seg'] * L['partial_blocks'], dtype=torch.float64, device=self.device)
            self.sums = torch.zeros(2 * L['max_seg'], dtype=torch.float64, device=self.device)
            self.flag = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.t_dev = torch.tensor([float(v) for v in t_host], dtype=torch.float64, device=self.device)
            self.out = torch.empty(self.T, self.L, dtype=self.dtype, device=self.device)
            if self.one_call:
                if len(_adjoint_bufs) > 4:
                    _adjoint_bufs.clear()
                _adjoint_bufs[(str(self.device), self.L, self.T)] = (self.bufs, self.ctl, self.partials, self.sums, self.flag,
                                                                     self.t_dev, self.out)
        self.seg_off = native.host_i64(self.offs)
        self.seg_len = native.host_i64(self.lens)
        nseg = len(y0)
        rt = list(rtol) if _is_iterable(rtol) else [rtol] * nseg
        at = list(atol) if _is_iterable(atol) else [atol] * nseg
        sharded = dist_state.group() is not None
        numel = native.host_i64([dist_state.global_numel(n, self.device) if (sharded and i < self.rep_from) else n
                                 for i, n in enumerate(self.lens)])
        self.ctl_args = (native.ptr(self.ctl), self.code, nseg, native.host_f64(rt), native.host_f64(at),
                         numel, _dflt(opts.get('safety', 0.9)), _dflt(opts.get('ifactor', 10.0)),
                         _dflt(opts.get('dfactor', 0.2)), _dflt(1 / 5), int(opts.get('max_num_steps', 2 ** 31 - 1)), self.T, 1)
        native.check(native.lib().node_b200_ctl_init(*self.ctl_args, native.stream_ptr()), 'ctl_init')
        self.y0_kept = y0 if self.one_call else None
        self.first_step = opts.get('first_step')
        key = 'ts32' if self.code == native.F32 else 'ts64'
        isz = 4 if self.code == native.F32 else 8
        self.ts = self.ctl[L[key]:L[key] + 7 * isz].view(self.dtype)
        for b, y in zip(self._views(0), y0):
            b.copy_(y.detach())

    def _views(self, i):
        row = self.bufs[i]
        return [row[o:o + n].view(s) for o, n, s in zip(self.offs, self.lens, self.shapes)]

    def _eval(self, t0d, src, dst):
        if hasattr(self.func, 'eval_into'):
            self.func.eval_into(t0d, self._views(src), self._views(dst), self.tsign)
        else:
            vals = self.func(t0d, tuple(self._views(src)))
            for b, v in zip(self._views(dst), vals):
                b.copy_(v)
        if dist_state.group() is not None and self.rep_from < len(self.lens):
            dist_state.all_reduce_sum(self.bufs[dst][self.offs[self.rep_from]:])   # partial batch sums -> global

    def _kptrs(self, idxs):
        arr = (native._vp * 7)()
        for j, i in enumerate(idxs):
            arr[j] = self.bufs[i].data_ptr()
        return arr

    def _reduce_and_control(self, mode):
        lib, sp = native.lib(), native.stream_ptr()
        nrows = 2 * len(self.lens)
        native.check(lib.node_b200_reduce_partials(native.ptr(self.partials), nrows, native.ptr(self.sums), sp), 'reduce')
        if dist_state.group() is not None:
            if self.rep_from < len(self.lens) and dist_state.rank() != 0:
                self.sums[2 * self.rep_from:].zero_()        # replicated members: counted once (rank 0's copy)
            if dist_state.peer_active():                     # in-kernel all-reduce over NVLink peer memory (csrc/peer_reduce.cu)
                native.check(lib.node_b200_fold_reduce(native.ptr(self.sums), 1, native.ptr(self.sums), len(self.sums), sp), 'fold_reduce')
            else:
                dist_state.all_reduce_sum(self.sums)
        native.check(lib.node_b200_controller(native.ptr(self.ctl), mode, native.ptr(self.sums),
                                              native.ptr(self.flag) if mode == 2 else native._vp(0),
                                              native.ptr(self.t_dev), sp), 'controller')

    def _run_one_call(self):
        """The whole interval by node_b200_adjoint_solve (csrc/adjoint_solve.cu): prologue + device-side while loop, ONE read of
        the controller block afterwards. Returns None when the loop graph could not be built (the caller takes the step-wise
        route from scratch)."""
        global _adjoint_solve_failed
        lib, sp = native.lib(), native.stream_ptr()
        N, C, H, W = self.shapes[0]
        fws = fused_workspace(self.device, N, C, H, W)
        fws.prepare(recognise_odefunc(self.func.func))
        vws = _vjp_workspace(self.device, N, C, H, W)
        # few images (PGD, evaluate.py nfe: SMs are idle beside k_vjp's handful of CTAs): a second VJP workspace lets the weight-gradient
        # GEMM of a stage run under the next stage's k_vjp. Measured 1-5 % of a batch-1 training step; nothing at batch 128, where
        # k_vjp's 128 one-image CTAs leave the GEMM no SM to run on.
        overlap = N <= 64 and os.environ.get('NODE_B200_ADJOINT_OVERLAP', '1') != '0'
        vws2 = _vjp_workspace(self.device, N, C, H, W, slot=1) if overlap else None
        given = self.first_step is not None
        if given:
            self.sums[0] = _dflt(0.01)                   # dopri5.py:81-82
        rc = lib.node_b200_adjoint_solve(native.ptr(self.ctl), native.ptr(self.bufs), self.L, self.seg_off, self.seg_len, 4,
                                         native.ptr(fws.buf), native.ptr(vws), native.ptr(vws2), float(self.tsign), native.layout()['ts32'],
                                         N, C, H, W,
                                         native.ptr(self.partials), native.ptr(self.sums), native.ptr(self.flag),
                                         native.ptr(self.t_dev), native.ptr(self.out), 1 if given else 0, sp)
        if rc != 0:
            torch.cuda.synchronize(self.device)
            _adjoint_solve_failed = True
            warnings.warn('node_b200_adjoint_solve: no device-side loop graph (CUDA error %d); the adjoint reads the controller '
                          'back once per attempted step instead' % rc)
            return None
        view = native.CtlView(self.ctl)                                       # the one host read of the interval
        if hasattr(self.func.target, 'nfe'):
            self.func.target.nfe += view.i32('nfe')                           # model.py:340 counts every evaluation
        last_stats.clear()
        last_stats.update(route='generic', nfe=view.i32('nfe'), n_accept=view.i32('n_accept'), n_reject=view.i32('n_reject'),
                          status=view.i32('status'), trace=view.trace(), adjoint_loop='device')
        native.raise_for_status(view.i32('status'))
        return tuple(self.out[:, o:o + n].reshape((self.T,) + s).clone() for o, n, s in zip(self.offs, self.lens, self.shapes))

    def run(self):
        if self.one_call:
            outs = self._run_one_call()
            if outs is not None:
                return outs
            for b, y in zip(self._views(0), self.y0_kept):                   # start over on the step-wise route
                b.copy_(y.detach())
            native.check(native.lib().node_b200_ctl_init(*self.ctl_args, native.stream_ptr()), 'ctl_init')
        lib, sp = native.lib(), native.stream_ptr()
        Y0, Y1, F0, F1, K2, YMID, YI = 0, 1, 2, 3, 4, 9, 10
        nseg = len(self.lens)
        segs = (self.seg_off, self.seg_len, nseg)
        ctl = native.ptr(self.ctl)
        t0d = torch.tensor(self.t_host[0], dtype=self.dtype, device=self.device)
        self._eval(t0d, Y0, F0)                                               # dopri5.py:78
        if self.first_step is None:
            native.check(lib.node_b200_init_norms(ctl, self.code, 0, native.ptr(self.bufs[Y0]), native.ptr(self.bufs[F0]),
                                                  native._vp(0), *segs, native.ptr(self.partials), sp), 'init_norms')
            self._reduce_and_control(0)
            native.check(lib.node_b200_rk_stage_combine(ctl, self.code, 7, native.ptr(self.bufs[YI]), native.ptr(self.bufs[Y0]),
                                                        self._kptrs([F0]), 1, self.L, sp), 'probe')
            self._eval(self.ts[1], YI, K2)                                    # misc.py:134
            native.check(lib.node_b200_init_norms(ctl, self.code, 1, native.ptr(self.bufs[Y0]), native.ptr(self.bufs[F0]),
                                                  native.ptr(self.bufs[K2]), *segs, native.ptr(self.partials), sp), 'init_norms')
            self._reduce_and_control(1)
        else:
            # dopri5.py:81-82: ANY non-None first_step means "start with 0.01" (built in the default dtype, then widened)
            self.sums[0] = _dflt(0.01)
            native.check(lib.node_b200_controller(ctl, 3, native.ptr(self.sums), native._vp(0), native.ptr(self.t_dev), sp),
                         'controller')
        self.out[0].copy_(self.bufs[Y0])
        cur = 0
        view = native.CtlView(self.ctl)
        # The adjoint of the fused ODE-Net dynamics on one GPU: a whole attempted step is ONE C call (node_b200_adjoint_step) -
        # the same kernels in the same order as the loop below, without ~20 Python -> C round trips per attempt.
        fast = None
        if (isinstance(self.func, _FusedAugmented) and dist_state.group() is None and self.code == native.F32 and nseg == 4
                and os.environ.get('NODE_B200_ADJOINT_STEP', '1') != '0'):
            N, C, H, W = self.shapes[0]
            fws = fused_workspace(self.device, N, C, H, W)
            fws.prepare(recognise_odefunc(self.func.func))
            fast = (fws.buf, _vjp_workspace(self.device, N, C, H, W), N, C, H, W)
        while not view.i32('done'):
            y, f, yn, fn = Y0 + cur, F0 + cur, Y0 + (cur ^ 1), F0 + (cur ^ 1)
            ks = [f, K2, K2 + 1, K2 + 2, K2 + 3, K2 + 4, fn]
            if fast is not None:
                native.check(lib.node_b200_adjoint_step(ctl, native.ptr(self.bufs), self.L, cur, self.seg_off, self.seg_len, nseg,
                                                        native.ptr(fast[0]), native.ptr(fast[1]), float(self.tsign),
                                                        native.layout()['ts32'], fast[2], fast[3], fast[4], fast[5],
                                                        native.ptr(self.partials), native.ptr(self.sums), native.ptr(self.flag),
                                                        native.ptr(self.t_dev), sp), 'adjoint_step')
                if hasattr(self.func.target, 'nfe'):
                    self.func.target.nfe += 6             # model.py:340 counts every evaluation
            else:
                for i in range(6):
                    dst = YI if i < 5 else yn
                    native.check(lib.node_b200_rk_stage_combine(ctl, self.code, i, native.ptr(self.bufs[dst]),
                                                                native.ptr(self.bufs[y]), self._kptrs(ks[:i + 1]), i + 1,
                                                                self.L, sp), 'stage_combine')
                    self._eval(self.ts[i + 1], dst, ks[i + 1])
                native.check(lib.node_b200_rk_error_norm(ctl, self.code, native.ptr(self.bufs[y]), native.ptr(self.bufs[yn]),
                                                         self._kptrs(ks), *segs, native.ptr(self.partials),
                                                         native.ptr(self.flag), sp), 'error_norm')
                self._reduce_and_control(2)
            view = native.CtlView(self.ctl)                                   # the one host read per attempt
            if view.i32('accepted_last'):
                if view.i32('out_hi') > view.i32('out_lo'):
                    native.check(lib.node_b200_rk_stage_combine(ctl, self.code, 6, native.ptr(self.bufs[YMID]),
                                                                native.ptr(self.bufs[y]), self._kptrs(ks), 7, self.L, sp), 'ymid')
                    native.check(lib.node_b200_interp_eval(ctl, self.code, native.ptr(self.t_dev), native.ptr(self.out), self.L,
                                                           native.ptr(self.bufs[y]), native.ptr(self.bufs[yn]),
                                                           native.ptr(self.bufs[YMID]), native.ptr(self.bufs[f]),
                                                           native.ptr(self.bufs[fn]), self.L, 0, sp), 'interp')
                cur ^= 1
        last_stats.clear()
        last_stats.update(route='generic', nfe=view.i32('nfe'), n_accept=view.i32('n_accept'),
                          n_reject=view.i32('n_reject'), status=view.i32('status'), trace=view.trace())
        native.raise_for_status(view.i32('status'))
        outs = []
        for o, n, s in zip(self.offs, self.lens, self.shapes):
            v = self.out[:, o:o + n].reshape((self.T,) + s)
            outs.append(v if (len(self.lens) == 1 and n == self.L) else v.contiguous())
        return tuple(outs)


def _dflt(x):
    """The reference builds its controller constants with `torch.tensor(x)` in the DEFAULT dtype and only
    then widens them to float64 (misc.py:37-44, dopri5.py:72-74, misc.py:168): keep that rounding."""
    return float(torch.tensor(x).type(torch.float64))


def _is_iterable(x):
    try:
        iter(x)
        return True
    except TypeError:
        return False


@native.on_device_of(1)
def _solve(func, y0, t, rtol, atol, options):
    """Shared by odeint and the adjoint's forward/backward: y0 is a tuple, returns a tuple."""
    for y in y0:
        if not y.is_cuda:
            raise RuntimeError('node_b200 is a CUDA-only implementation of the dopri5 hot path: the state must live '
                               'on a B200 (got a %s tensor). There is no CPU fallback.' % y.device)
    t_host = _host_times(t)
    if len(t_host) > 1 and bool((t_host[1:] < t_host[:-1]).all()):            # misc.py:184-187
        tsign, t_host = -1, -t_host
    else:
        tsign = 1
    assert bool((t_host[1:] > t_host[:-1]).all()), 't must be strictly increasing or decrasing'
    unknown = [k for k in options if k not in _DOPRI5_OPTIONS]
    if unknown:
        warnings.warn('Dopri5Solver: Unexpected arguments {}'.format({k: options[k] for k in unknown}))
    with torch.no_grad(), _nvtx('node_b200.solve'):
        params = recognise_odefunc(func)
        plain = not options and not _is_iterable(rtol) and not _is_iterable(atol)
        if params is not None and plain and _fusable_state(params, y0) and len(t_host) <= 1024:
            return (_solve_fused(func, params, y0[0], t_host, tsign, rtol, atol),)
        from . import wide
        base_mod = _unwrap(func)
        if (plain and os.environ.get('NODE_B200_WIDE', '1') != '0' and os.environ.get('NODE_B200_FUSED', '1') != '0'
                and wide.serves(base_mod, y0)):
            # n_filters = 128 / 192 / 256 (the paper's CIFAR setting): the dynamics as 64-channel blocks on the tcgen05 engine,
            # K2-K6 from the generic route
            out = _GenericSolve(wide.WideDynamics.of(base_mod), y0, t_host, rtol, atol, options, tsign=tsign).run()
            last_stats['route'] = 'native-wide'
            return out
        if tsign < 0 and not hasattr(func, 'eval_into'):
            base = func
            call = lambda tt, yy: tuple(-v for v in base(-tt, yy))
        else:
            call = func
        return _GenericSolve(call, y0, t_host, rtol, atol, options, tsign=tsign,
                             rep_from=getattr(func, 'replicated_from', None)).run()


def odeint(func, y0, t, rtol=1e-7, atol=1e-9, method=None, options=None):
    """Integrate dy/dt = func(t, y), y(t[0]) = y0 and return y at every t (odeint.py:20-76).

    Same contract as the reference: y0 a Tensor or tuple of Tensors, t 1-D strictly monotone,
    result [len(t), *y0.shape] with result[0] == y0. Only method='dopri5' (the default) is served.
    """
    tensor_input, y0 = _check_inputs(func, y0, t)
    if options is None:
        options = {}
    elif method is None:
        raise ValueError('cannot supply `options` without specifying `method`')
    if method is None:
        method = 'dopri5'
    if method not in _KNOWN_METHODS:
        raise KeyError(method)
    if method != 'dopri5':
        raise NotImplementedError("node_b200 implements method='dopri5' only (got %r)" % method)
    if _needs_grad(func, y0, t):
        # The reference unrolls autograd through every solver op (train.py without --adjoint, model.py:359). The solver kernels
        # record no graph, so this request is served by node_b200.unrolled: the reference's solver loop recorded by autograd
        # (controller terms included - the same gradient), every evaluation of the recognised dynamics and its VJP on the native
        # kernels. NODE_B200_ODEINT_GRAD=adjoint serves it by the adjoint ODE instead (O(1) memory, a different gradient that
        # agrees to the solver tolerance: gradient_tests.py:98-116).
        if os.environ.get('NODE_B200_ODEINT_GRAD', 'unrolled') != 'adjoint':
            for y in y0:
                if not y.is_cuda:
                    raise RuntimeError('node_b200 is a CUDA-only implementation of the dopri5 hot path: the state must live '
                                       'on a B200 (got a %s tensor). There is no CPU fallback.' % y.device)
            unknown = [k for k in options if k not in ('first_step', 'safety', 'ifactor', 'dfactor', 'max_num_steps')]
            if unknown:
                warnings.warn('Dopri5Solver: Unexpected arguments {}'.format({k: options[k] for k in unknown}))
            from . import unrolled
            user = func
            call = (lambda tt, yy: (user(tt, yy[0]),)) if tensor_input else func
            if tensor_input and isinstance(user, nn.Module):
                call = _TensorFunc(user)
            params = recognise_odefunc(user) if (tensor_input and isinstance(user, nn.Module)) else None
            if (params is not None and not options and not _is_iterable(rtol) and not _is_iterable(atol) and _fusable_state(params, y0)
                    and native.lib().node_b200_vjp_workspace_bytes(*[int(v) for v in y0[0].shape]) > 0
                    and os.environ.get('NODE_B200_ODEINT_GRAD', 'unrolled') != 'unrolled-eager'):
                # The recognised dynamics: the forward is the fused solve (no graph, O(1) memory, full speed - evaluate.py:109 calls
                # the model with gradients enabled and never calls backward); the unrolled graph is recorded only if a gradient is
                # actually asked for (unrolled.LazyUnrolled: replay under autograd inside backward).
                return unrolled.LazyUnrolled.apply(user, call, t, rtol, atol, y0[0], *tuple(user.parameters()))
            last_stats.clear()
            with torch.cuda.device(y0[0].device), _nvtx('node_b200.unrolled'):
                out = unrolled.solve(call, y0, t, rtol, atol, options, stats=last_stats)
            return out[0] if tensor_input else out
        if not isinstance(func, nn.Module):
            if not callable(func):
                raise NotImplementedError('node_b200.odeint: gradients need a callable `func`')
            func = _CallableModule(func)             # parameters of every module the callable closes over get their gradient
        global _warned_grad_bridge
        if not _warned_grad_bridge:
            warnings.warn('node_b200.odeint: gradients requested through the non-adjoint entry point - served by the '
                          'adjoint ODE (odeint_adjoint) at the same rtol/atol')
            _warned_grad_bridge = True
        return odeint_adjoint(func, y0[0] if tensor_input else y0, t, rtol=rtol, atol=atol, method=method, options=options or None)
    if tensor_input:
        user = func
        if isinstance(user, nn.Module):
            func = _TensorFunc(user)
        else:
            func = lambda tt, yy: (user(tt, yy[0]),)
        return _solve(func, y0, t, rtol, atol, options)[0]
    return _solve(func, y0, t, rtol, atol, options)


# ---- adjoint (adjoint.py:7-133) --------------------------------------------------------------------

def _flatten(seq):
    flat = [p.contiguous().view(-1) for p in seq]
    return torch.cat(flat) if flat else torch.tensor([])


class _AdjointFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, func, t, rtol, atol, options, n_tensors, flat_params, *y0):
        ctx.func, ctx.rtol, ctx.atol, ctx.options, ctx.n = func, rtol, atol, options, n_tensors
        with torch.no_grad():
            ans = _solve(func, tuple(y0), t, rtol, atol, options)
        ctx.save_for_backward(t, flat_params, *ans)
        return ans

    @staticmethod
    def backward(ctx, *grad_output):
        t, flat_params, *ans = ctx.saved_tensors
        func, rtol, atol, options, n = ctx.func, ctx.rtol, ctx.atol, ctx.options, ctx.n
        f_params = tuple(func.parameters())
        t_host = _host_times(t)

        def augmented(tt, y_aug):                                             # adjoint.py:32-55
            y, adj_y = y_aug[:n], y_aug[n:2 * n]
            with torch.enable_grad():
                tt = tt.detach().clone().requires_grad_(True)
                y = tuple(v.detach().clone().requires_grad_(True) for v in y)
                fe = func(tt, y)
                g = torch.autograd.grad(fe, (tt,) + y + f_params, tuple(-a for a in adj_y), allow_unused=True)
            vt = torch.zeros_like(tt) if g[0] is None else g[0]
            vy = tuple(torch.zeros_like(a) if b is None else b for a, b in zip(y, g[1:1 + n]))
            vp = [torch.zeros_like(p).view(-1) if q is None else q.contiguous().view(-1)
                  for q, p in zip(g[1 + n:], f_params)]
            vp = torch.cat(vp) if vp else torch.zeros(1, dtype=vy[0].dtype, device=vy[0].device).view(())
            return tuple(v.detach() for v in fe) + vy + (vt.reshape(()), vp)

        T = ans[0].shape[0]
        native_vjp = (n == 1 and os.environ.get('NODE_B200_NATIVE_VJP', '1') != '0' and recognise_odefunc(func) is not None
                      and _fusable_state(recognise_odefunc(func), (ans[0][0],))
                      and native.lib().node_b200_vjp_workspace_bytes(*[int(v) for v in ans[0].shape[1:]]) > 0)
        from . import wide
        wide_vjp = (not native_vjp and n == 1 and os.environ.get('NODE_B200_NATIVE_VJP', '1') != '0'
                    and os.environ.get('NODE_B200_WIDE', '1') != '0' and os.environ.get('NODE_B200_FUSED', '1') != '0'
                    and wide.serves(_unwrap(func), (ans[0][0],)))
        if native_vjp:
            augmented = _FusedAugmented(func)
        elif wide_vjp:
            augmented = wide.WideAugmented(func)    # n_filters = 128 / 192 / 256 (reproduce.sh:21-25 trains these with --adjoint)
        else:
            augmented.replicated_from = 2 * n       # adj_t, adj_params: sums over the (possibly sharded) batch
        with torch.no_grad(), _nvtx('node_b200.adjoint_backward'):
            adj_y = tuple(g[-1] for g in grad_output)
            adj_p = torch.zeros_like(flat_params)
            adj_t = torch.zeros((), dtype=ans[0].dtype, device=ans[0].device)
            tv = []
            for i in range(T - 1, 0, -1):
                ans_i = tuple(a[i] for a in ans)
                ti = torch.tensor(t_host[i], dtype=ans[0].dtype, device=ans[0].device)
                if native_vjp:
                    func_i = (odefunc_forward(func, float(t_host[i]), ans_i[0].contiguous()),)
                    _unwrap(func).nfe += 1
                elif wide_vjp:
                    func_i = (augmented.dyn(float(t_host[i]), ans_i[0].contiguous()),)      # counts itself
                else:
                    func_i = func(ti, ans_i)
                d = sum(torch.dot(f.reshape(-1), g[i].reshape(-1)).view(1) for f, g in zip(func_i, grad_output))
                if dist_state.group() is not None:
                    dist_state.all_reduce_sum(d)              # dL/dt is a sum over the whole (sharded) batch
                adj_t = adj_t - d.reshape(())
                tv.append(d)
                if adj_p.numel() == 0:
                    adj_p = torch.zeros((), dtype=adj_y[0].dtype, device=adj_y[0].device)
                aug0 = ans_i + tuple(a.contiguous() for a in adj_y) + (adj_t, adj_p)
                span = torch.tensor([t_host[i], t_host[i - 1]], dtype=torch.float64)
                sol = _solve(augmented, aug0, span, rtol, atol, options)
                adj_y = tuple(s[1] for s in sol[n:2 * n])
                adj_t = sol[2 * n][1]
                adj_p = sol[2 * n + 1][1]
                adj_y = tuple(a + g[i - 1] for a, g in zip(adj_y, grad_output))
            tv.append(adj_t.reshape(1))
            time_vjps = torch.cat(tv[::-1]).to(t.dtype)
        last_stats['adjoint_vjp'] = 'native' if native_vjp else ('native-wide' if wide_vjp else 'autograd')
        return (None, time_vjps, None, None, None, None, adj_p) + tuple(adj_y)


def odeint_adjoint(func, y0, t, rtol=1e-6, atol=1e-12, method=None, options=None):
    """odeint with O(1)-memory gradients via the adjoint ODE (adjoint.py:105-133)."""
    if not isinstance(func, nn.Module):
        raise ValueError('func is required to be an instance of nn.Module.')
    tensor_input, y0 = _check_inputs(func, y0, t)
    if options is None:
        options = {}
    elif method is None:
        raise ValueError('cannot supply `options` without specifying `method`')
    if method is not None and method not in _KNOWN_METHODS:
        raise KeyError(method)
    if method not in (None, 'dopri5'):
        raise NotImplementedError("node_b200 implements method='dopri5' only (got %r)" % method)
    if tensor_input:
        func = _TensorFunc(func)
    flat_params = _flatten(func.parameters())
    ys = _AdjointFn.apply(func, t, rtol, atol, options, len(y0), flat_params, *y0)
    return ys[0] if tensor_input else ys
