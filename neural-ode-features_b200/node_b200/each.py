"""Batched INDEPENDENT solvers: N dopri5 solves with per-sample error norms, step sizes and accept / reject sequences
(SURVEY 8f-4). The reference gets them by running the model with batch_size = 1 - `evaluate.py nfe` (evaluate.py:109-126) records
the NFE of every test image that way, and the adversarial attacks (adversarial/attack.py:63-72) integrate single images - one
solve after the other, each bound by its own launch latency.

Here every sample is still its own solve on the fused route (so the result is the one `odeint(func, y0[i:i+1], ...)` returns:
same kernels, same controller), but K solves are in flight at a time: K private workspaces, K CUDA streams, one captured
launch sequence per workspace (f0, initial-step probe, `steps` attempted steps with their controllers and dense output, all
no-ops once the controller says done), and the controller blocks come back through pinned memory with ONE synchronisation at
the end. A batch-1 solve occupies one SM, so K of them overlap on the device.
"""
import os

import torch

from . import native
from . import solver as _solver


class _Lane(object):
    """One in-flight solve: private workspace, stream, captured launch sequence."""

    def __init__(self, device, shape, T):
        C, H, W = shape
        self.ws = _solver.FusedWorkspace(device, 1, C, H, W)
        self.stream = torch.cuda.Stream(device=device)
        self.y = torch.empty((1, C, H, W), dtype=torch.float32, device=device)
        self.out = torch.empty((T, 1, C, H, W), dtype=torch.float32, device=device)
        self.graph = None
        self.key = None


_lanes = {}


def _get_lanes(device, shape, T, K):
    key = (str(device), tuple(shape), T)
    lanes = _lanes.get(key)
    if lanes is None or len(lanes) < K:
        if len(_lanes) > 4:
            _lanes.clear()
        lanes = _lanes[key] = [_Lane(device, shape, T) for _ in range(K)]
    return lanes[:K]


@native.on_device_of(1)
def odeint_each(func, y0, t, rtol=1e-7, atol=1e-9, method=None, lanes=16, steps=None):
    """out[T, N, ...], stats: the i-th column equals odeint(func, y0[i:i+1], t, rtol, atol)[:, 0]; stats[i] = dict(nfe, n_accept,
    n_reject). `func` must be the ODE-Net dynamics served by the fused route (fp32 CUDA, 64 filters); `func.nfe` grows by the
    total. `steps` = attempted steps enqueued per solve (default: a running estimate); a solve that needs more is finished by the
    ordinary batch-1 route."""
    if method not in (None, 'dopri5'):
        raise NotImplementedError("node_b200 implements method='dopri5' only (got %r)" % method)
    tensor_input, y0t = _solver._check_inputs(func, y0, t)
    if not tensor_input:
        raise TypeError('odeint_each takes a tensor state [N, C, H, W]')
    y = y0t[0]
    params = _solver.recognise_odefunc(func)
    if params is None or y.dim() != 4 or not _solver._fusable_state(params, (y[:1],)):
        raise NotImplementedError('odeint_each serves the fused ODE-Net route (fp32 CUDA state [N, 64, H, W]) only')
    t_host, tsign = _solver._host_times(t), 1
    if len(t_host) > 1 and bool((t_host[1:] < t_host[:-1]).all()):     # misc.py:184-187
        t_host, tsign = -t_host, -1
    assert bool((t_host[1:] > t_host[:-1]).all()), 't must be strictly increasing or decrasing'
    y = y.detach().contiguous()
    N, C, H, W = (int(v) for v in y.shape)
    T = len(t_host)
    K = max(1, min(int(lanes), N))
    dev = y.device
    lib = native.lib()
    L = native.layout()
    conv_mode = _solver._conv_mode()
    th = native.host_f64(t_host)
    common = (th, T, float(rtol), float(atol), 1, C, H, W)
    E = C * H * W
    target = _solver._unwrap(func)
    if steps is None:
        steps = _solver._step_guess.get(('each', target), 8)
    steps = int(steps)
    use_graph = os.environ.get('NODE_B200_GRAPH', 'auto') != '0'
    out = torch.empty((T, N, C, H, W), dtype=torch.float32, device=dev)
    ctl_host = torch.empty((N, L['sizeof']), dtype=torch.uint8).pin_memory()
    ls = _get_lanes(dev, (C, H, W), T, K)
    main = torch.cuda.current_stream(dev)
    ready = torch.cuda.Event()
    ready.record(main)
    gkey = (tuple(float(v) for v in t_host), float(rtol), float(atol), conv_mode, tsign, steps)
    for lane in ls:
        lane.stream.wait_event(ready)
        with torch.cuda.stream(lane.stream):
            lane.ws.prepare(params)
            if use_graph and (lane.graph is None or lane.key != gkey):
                lane.graph, lane.key = torch.cuda.CUDAGraph(), gkey
                try:
                    with torch.cuda.graph(lane.graph, stream=lane.stream):
                        native.check(lib.node_b200_fused_solve(native.ptr(lane.ws.buf), native.ptr(lane.y), *common, E, native.ptr(lane.out),
                                                               conv_mode, tsign, 1, steps, native.stream_ptr()), 'fused_solve (capture)')
                except Exception:
                    lane.graph, use_graph = None, False
    for i in range(N):
        lane = ls[i % K]
        with torch.cuda.stream(lane.stream):
            lane.y.copy_(y[i:i + 1], non_blocking=True)
            if use_graph and lane.graph is not None:
                lane.graph.replay()
            else:
                native.check(lib.node_b200_fused_solve(native.ptr(lane.ws.buf), native.ptr(lane.y), *common, E, native.ptr(lane.out),
                                                       conv_mode, tsign, 1, steps, native.stream_ptr()), 'fused_solve')
            out[:, i:i + 1].copy_(lane.out, non_blocking=True)
            ctl_host[i].copy_(lane.ws.ctl, non_blocking=True)
    for lane in ls:
        main.wait_stream(lane.stream)
    torch.cuda.synchronize(dev)
    stats, most = [], 0
    raw = ctl_host.numpy()

    def i32(row, name):
        return int(raw[row, L[name]:L[name] + 4].view('int32')[0])

    total = 0
    for i in range(N):
        if not i32(i, 'done'):
            # more attempted steps than were enqueued: the ordinary batch-1 route finishes this sample from scratch
            nfe0 = getattr(target, 'nfe', 0)
            o = _solver._solve_fused(func, params, y[i:i + 1], t_host, tsign, rtol, atol)
            if hasattr(target, 'nfe'):
                target.nfe = nfe0
            out[:, i:i + 1].copy_(o)
            st = dict(nfe=_solver.last_stats['nfe'], n_accept=_solver.last_stats['n_accept'], n_reject=_solver.last_stats['n_reject'])
        else:
            native.raise_for_status(i32(i, 'status'))
            st = dict(nfe=i32(i, 'nfe'), n_accept=i32(i, 'n_accept'), n_reject=i32(i, 'n_reject'))
        most = max(most, st['n_accept'] + st['n_reject'])
        total += st['nfe']
        stats.append(st)
    _solver._step_guess[('each', target)] = most + 1
    if hasattr(target, 'nfe'):
        target.nfe += total
    _solver.last_stats.clear()
    _solver.last_stats.update(route='fused-each', nfe=total, lanes=K, steps_enqueued=steps)
    return out, stats
