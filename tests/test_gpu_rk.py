"""GPU: K2-K6 building blocks and the generic-callable route against the oracle and the golden
vectors recorded from the reference. Bit-exact where the reference is elementwise; reductions and
libm calls (sum order, powf) to 1e-6."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import dopri5_port
from test_oracle import make_problem

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _ctl(native, dtype, rtol, atol, numels, n_out=2):
    L = native.layout()
    ctl = torch.zeros(L['sizeof'], dtype=torch.uint8, device=DEV)
    code = native.F32 if dtype == torch.float32 else native.F64
    n = len(numels)
    err = native.lib().node_b200_ctl_init(
        native.ptr(ctl), code, n, native.host_f64([rtol] * n), native.host_f64([atol] * n),
        native.host_i64(numels), float(np.float32(0.9)), 10.0, float(np.float32(0.2)), float(np.float32(0.2)),
        2 ** 31 - 1, n_out, 1, native.stream_ptr())
    native.check(err, 'ctl_init')
    return ctl, code


def _poke(ctl, native, name, value, np_dtype):
    off = native.layout()[name]
    raw = np.array([value], dtype=np_dtype).tobytes()
    ctl[off:off + len(raw)] = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(DEV)


def _kptrs(ts):
    arr = (ctypes.c_void_p * 7)()
    for i, t in enumerate(ts):
        arr[i] = t.data_ptr()
    return arr


@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
@pytest.mark.parametrize('numel', [8, 1000, 262144 + 3])
def test_stage_combine_and_error_norm_match_reference_arithmetic(native_lib, dtype, numel):
    from node_b200 import native
    torch.manual_seed(numel)
    y0 = torch.randn(numel, dtype=dtype)
    ks = [torch.randn(numel, dtype=dtype) for _ in range(7)]
    h64 = 0.0817
    h = torch.tensor(h64, dtype=torch.float64).to(dtype)
    ctl, code = _ctl(native, dtype, 1e-3, 1e-3, [numel])
    _poke(ctl, native, 'h32', float(h), np.float32)
    _poke(ctl, native, 'h64', float(h), np.float64)
    _poke(ctl, native, 'it_h32', float(h), np.float32)
    _poke(ctl, native, 'it_h64', float(h), np.float64)
    _poke(ctl, native, 'out_hi', 1, np.int32)            # row 6 (mid-point) returns at once unless the controller scheduled outputs
    dy0, dks = y0.to(DEV), [k.to(DEV) for k in ks]
    out = torch.empty_like(dy0)
    for row in range(6):
        err = native_lib.node_b200_rk_stage_combine(native.ptr(ctl), code, row, native.ptr(out), native.ptr(dy0),
                                                   _kptrs(dks[:row + 1]), row + 1, numel, native.stream_ptr())
        native.check(err, 'combine')
        ref = y0 + dopri5_port.weighted_sum(h, dopri5_port.BETA[row], ks[:row + 1])
        assert torch.equal(out.cpu(), ref), 'stage %d not bit-exact' % (row + 1)
    native.check(native_lib.node_b200_rk_stage_combine(native.ptr(ctl), code, 6, native.ptr(out), native.ptr(dy0), _kptrs(dks), 7,
                                                       numel, native.stream_ptr()), 'ymid')
    assert torch.equal(out.cpu(), y0 + dopri5_port.weighted_sum(h, dopri5_port.C_MID, ks))
    # error norm: y1 = stage-6 combination
    y1 = y0 + dopri5_port.weighted_sum(h, dopri5_port.BETA[5], ks[:6])
    L = native.layout()
    partials = torch.zeros(2 * L['max_seg'] * L['partial_blocks'], dtype=torch.float64, device=DEV)
    sums = torch.zeros(2 * L['max_seg'], dtype=torch.float64, device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    err = native_lib.node_b200_rk_error_norm(native.ptr(ctl), code, native.ptr(dy0), native.ptr(y1.to(DEV)), _kptrs(dks),
                                            native.host_i64([0]), native.host_i64([numel]), 1,
                                            native.ptr(partials), native.ptr(flag), native.stream_ptr())
    native.check(err, 'error_norm')
    native.check(native_lib.node_b200_reduce_partials(native.ptr(partials), 2, native.ptr(sums), native.stream_ptr()), 'reduce')
    ratio = dopri5_port.error_ratios([dopri5_port.weighted_sum(h, dopri5_port.C_ERR, ks)], 1e-3, 1e-3, [y0], [y1])[0]
    got = float(sums[0]) / numel
    assert abs(got - float(ratio)) <= 2e-6 * float(ratio)
    assert int(flag) == 0
    dy0[numel // 2] = float('inf')
    native_lib.node_b200_rk_error_norm(native.ptr(ctl), code, native.ptr(dy0), native.ptr(y1.to(DEV)), _kptrs(dks),
                                       native.host_i64([0]), native.host_i64([numel]), 1,
                                       native.ptr(partials), native.ptr(flag), native.stream_ptr())
    assert int(flag) == 1                                          # dopri5.py:102


def test_controller_step_size_rule(native_lib):
    """misc.py:160-170 on the device, incl. the float32-rounded constants of dopri5.py:72-74."""
    from node_b200 import native
    L = native.layout()
    for ratio in (0.0, 1e-6, 0.3, 0.999, 1.0, 1.0001, 2.5, 1e4):
        ctl, code = _ctl(native, torch.float32, 1e-3, 1e-3, [1000], n_out=2)
        t_out = torch.tensor([0.0, 100.0], dtype=torch.float64, device=DEV)
        sums = torch.zeros(16, dtype=torch.float64, device=DEV)
        # INIT_B with d1 = d2 = 1 gives dt0 = min(100*h0, (0.01)^(1/5)); then one STEP with the chosen ratio
        _poke(ctl, native, 'h0', 0.001, np.float64)
        sums[0] = 1000.0 * (0.001 ** 2)
        native.check(native_lib.node_b200_controller(native.ptr(ctl), 1, native.ptr(sums), None, native.ptr(t_out), native.stream_ptr()), 'c1')
        v0 = native.CtlView(ctl)
        dt0 = v0.f64('dt')
        sums[0] = ratio * 1000.0
        flag = torch.zeros(1, dtype=torch.int32, device=DEV)
        native.check(native_lib.node_b200_controller(native.ptr(ctl), 2, native.ptr(sums), native.ptr(flag), native.ptr(t_out), native.stream_ptr()), 'c2')
        v = native.CtlView(ctl)
        r32 = torch.tensor(ratio * 1000.0, dtype=torch.float64).div(1000).to(torch.float32)
        want = float(dopri5_port.next_step_size(torch.tensor(dt0, dtype=torch.float64), [r32]))
        assert v.i32('accepted_last') == int(float(r32) <= 1.0)
        assert abs(v.f64('dt') - want) <= 1e-12 * want, (ratio, v.f64('dt'), want)
        assert v.i32('nfe') == 6 and v.i32('n_attempt') == 1
        if float(r32) <= 1:
            assert v.f64('t1') == dt0 and v.f64('t0') == 0.0
        else:
            assert v.f64('t1') == 0.0


def _run_generic(f, y0, t, rtol=1e-7, atol=1e-9, **kw):
    from node_b200 import odeint, solver
    out = odeint(f, y0, t, rtol=rtol, atol=atol, **kw)
    return out, dict(solver.last_stats)


@pytest.fixture
def default_f64():
    """The reference's own tests run with torch.set_default_dtype(float64) (odeint_tests.py:9); the
    controller constants are built in the default dtype (dopri5.py:72-74), so the golden traces were too."""
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(prev)


@pytest.mark.parametrize('ode', ['constant', 'linear', 'sine'])
@pytest.mark.parametrize('direction', ['fwd', 'rev'])
def test_generic_route_f64_matches_reference_outputs(native_lib, golden, default_f64, ode, direction):
    """The reference's own analytic problems (torchdiffeq/tests/problems.py, odeint_tests.py:54-67,104-118)."""
    g = golden('generic_f64')
    key = '%s_%s' % (ode, direction)
    y0, t = torch.from_numpy(g[key + '.y0']).to(DEV), torch.from_numpy(g[key + '.t']).to(DEV)
    f = make_problem(ode, g.get(key + '.A'))
    out, st = _run_generic(f, y0, t)
    assert st['route'] == 'generic' and st['status'] == 0
    ref = torch.from_numpy(g[key + '.out'])
    assert float((out.cpu() - ref).abs().max() / ref.abs().max()) < 1e-9
    assert list(st['trace']['accepted']) == list(g[key + '.tr_acc'])
    np.testing.assert_allclose(st['trace']['dt'], g[key + '.tr_dt'], rtol=1e-9)
    assert st['nfe'] == 2 + 6 * len(g[key + '.tr_acc'])
    exact = torch.from_numpy(g[key + '.exact']).reshape(ref.shape)
    assert float((out.cpu() - exact).abs().max() / exact.abs().max()) < 1e-4
    assert torch.equal(out[0], y0)


def test_generic_route_tuple_state_and_single_time_point(native_lib, golden, default_f64):
    g = golden('generic_f64')
    A = torch.from_numpy(g['tuple.A']).to(DEV)
    y0, z0, t = (torch.from_numpy(g['tuple.' + k]).to(DEV) for k in ('y0', 'z0', 't'))
    f = lambda tt, y: (torch.mm(A, y[0].reshape(-1, 1)).reshape(-1), -0.5 * y[1])
    out, st = _run_generic(f, (y0, z0), t)                        # api_tests.py:19-38
    assert isinstance(out, tuple) and out[1].shape == (len(t), 3, 4)
    for o, k in zip(out, ('out0', 'out1')):
        ref = torch.from_numpy(g['tuple.' + k])
        assert float((o.cpu() - ref).abs().max() / ref.abs().max()) < 1e-9
    one, st = _run_generic(f, (y0, z0), t[:1])                    # odeint_tests.py:121-151: no integration
    assert one[0].shape == (1,) + tuple(y0.shape) and torch.equal(one[0][0], y0) and st['nfe'] == 2


def test_generic_route_f32_odefunc_callable(native_lib, golden):
    """Unfused float32 route with the eager dynamics as the callable: same decisions as the reference."""
    from conftest import load_odefunc
    g = golden('cifar_res_n8')
    func = load_odefunc(g, DEV)
    func.__class__ = type('PlainDynamics', (func.__class__,), {})   # defeat the recogniser by name
    h0, t = torch.from_numpy(g['h0']).to(DEV), torch.from_numpy(g['t']).to(DEV)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            out, st = _run_generic(func, h0, t, rtol=1e-3, atol=1e-3, method='dopri5')
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert st['route'] == 'generic'
    ref = torch.from_numpy(g['out'])
    assert float((out.cpu() - ref).abs().max() / ref.abs().max()) < 1e-4
    assert list(st['trace']['accepted']) == list(g['tr_acc']) and func.nfe == int(g['nfe'])
    np.testing.assert_allclose(st['trace']['dt'], g['tr_dt'], rtol=1e-5)


def test_dt_underflow_and_nonfinite_raise_assertion(native_lib):
    from node_b200 import odeint
    y0 = torch.ones(16, dtype=torch.float64, device=DEV)
    t = torch.tensor([0., 1.], dtype=torch.float64, device=DEV)
    with pytest.raises(AssertionError):
        odeint(lambda tt, y: y * float('nan'), y0, t)
    bad = y0.clone()
    bad[3] = float('inf')
    with pytest.raises(AssertionError):
        odeint(lambda tt, y: -y, bad, t)                           # non-finite state
