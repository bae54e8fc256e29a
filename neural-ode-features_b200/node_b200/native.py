"""ctypes binding of libnode_b200.so (the C ABI declared in include/node_b200.h).

There is no fallback: if the shared library is missing, stale (ABI mismatch) or the tensors are
not on a CUDA device, the solver raises.  Build with `python __graft_entry__.py build` (or
`python -c "import __graft_entry__ as g; g.build()"`) - nvcc cross-compiles for sm_100a.
"""
import ctypes
import functools
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('NODE_B200_LIB') or os.path.join(_HERE, 'lib', 'libnode_b200.so')   # override: tuning builds
ABI_VERSION = 1

F32, F64 = 0, 1
ST_DT_UNDERFLOW, ST_NONFINITE, ST_MAX_STEPS, ST_INTERP_RANGE, ST_WATCHDOG = 1, 2, 4, 8, 16

# order of the values returned by node_b200_ctl_layout (csrc/rk_kernels.cu)
CTL_FIELDS = (
    'sizeof', 'partial_blocks', 'max_seg', 'max_trace',
    't0', 't1', 'dt', 'ratio', 'ts64', 'ts32', 'h64', 'h32', 'out_lo', 'out_hi', 'next_out',
    'n_attempt', 'n_accept', 'n_reject', 'nfe', 'status', 'done', 'cur', 'accepted_last',
    'tr_t', 'tr_dt', 'tr_ratio', 'tr_acc', 'h0', 'h0_32', 'it_t0', 'it_t1', 'it_h64', 'it_h32',
)

EXPORTS = (
    'node_b200_abi_version', 'node_b200_ctl_layout', 'node_b200_ctl_init', 'node_b200_rk_stage_combine',
    'node_b200_rk_error_norm', 'node_b200_init_norms', 'node_b200_reduce_partials', 'node_b200_controller',
    'node_b200_interp_eval', 'node_b200_fused_workspace_bytes', 'node_b200_fused_prepare',
    'node_b200_odefunc_forward', 'node_b200_fused_solve', 'node_b200_fused_phase', 'node_b200_fused_sums',
    'node_b200_fused_ctl', 'node_b200_vjp_workspace_bytes', 'node_b200_odefunc_vjp', 'node_b200_wgrad',
    'node_b200_vjp_buffer', 'node_b200_groupnorm_relu', 'node_b200_resconv_workspace_bytes', 'node_b200_resconv_prepare',
    'node_b200_resconv_forward', 'node_b200_convs2_workspace_bytes', 'node_b200_convs2_prepare', 'node_b200_convs2_forward',
    'node_b200_stem_gn_relu', 'node_b200_head', 'node_b200_feature_normalize', 'node_b200_retrieval_scores',
    'node_b200_groupnorm_relu_backward', 'node_b200_absmax', 'node_b200_plane_split', 'node_b200_plane_merge',
    'node_b200_conv3x3_prepare', 'node_b200_conv3x3_forward', 'node_b200_conv_wgrad_workspace_bytes', 'node_b200_conv_wgrad',
    'node_b200_stem_backward_workspace_bytes', 'node_b200_stem_backward', 'node_b200_resconv_scal_offset',
    'node_b200_convs2_scal_offset', 'node_b200_peer_alloc', 'node_b200_peer_open', 'node_b200_peer_close', 'node_b200_peer_world',
    'node_b200_fold_reduce', 'node_b200_adjoint_step', 'node_b200_conv3x3_forward_strided', 'node_b200_groupnorm_relu_ex',
    'node_b200_wide_odefunc', 'node_b200_wide8_workspace_bytes', 'node_b200_wide8_operand_bytes', 'node_b200_wide8_prepare',
    'node_b200_wide8_gn_operand', 'node_b200_wide8_conv', 'node_b200_wide8_watchdog', 'node_b200_wide8_odefunc',
    'node_b200_adjoint_solve', 'node_b200_adjoint_solve_reset', 'node_b200_groupnorm_backward_ex', 'node_b200_batch_colsum',
    'node_b200_pow2_scale', 'node_b200_wide_conv_blocks', 'node_b200_wide_vjp', 'node_b200_wide8_raw_operand',
    'node_b200_lincomb', 'node_b200_lincomb_scale', 'node_b200_lincomb_dots', 'node_b200_lincomb_scratch_doubles',
    'node_b200_odefunc_vjp_split',
)

_lib = None
_layout = None

_vp, _i, _i64, _d, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_float


def _declare(lib):
    lib.node_b200_abi_version.restype = _i
    lib.node_b200_ctl_layout.argtypes = [_vp, _i]
    lib.node_b200_ctl_init.argtypes = [_vp, _i, _i, _vp, _vp, _vp, _d, _d, _d, _d, _i, _i, _i, _vp]
    lib.node_b200_rk_stage_combine.argtypes = [_vp, _i, _i, _vp, _vp, _vp, _i, _i64, _vp]
    lib.node_b200_rk_error_norm.argtypes = [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]
    lib.node_b200_init_norms.argtypes = [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]
    lib.node_b200_reduce_partials.argtypes = [_vp, _i, _vp, _vp]
    lib.node_b200_controller.argtypes = [_vp, _i, _vp, _vp, _vp, _vp]
    lib.node_b200_interp_eval.argtypes = [_vp, _i, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i, _vp]
    lib.node_b200_fused_workspace_bytes.argtypes = [_i, _i, _i, _i]
    lib.node_b200_fused_workspace_bytes.restype = _i64
    lib.node_b200_fused_prepare.argtypes = [_vp, _i, _i, _i] + [_vp] * 10 + [_f, _vp]
    lib.node_b200_odefunc_forward.argtypes = [_vp, _vp, _f, _f, _vp, _i, _i, _i, _i, _i, _vp]
    lib.node_b200_fused_solve.argtypes = [_vp, _vp, _vp, _i, _d, _d, _i, _i, _i, _i, _i64, _vp, _i, _i, _i, _i, _vp]
    lib.node_b200_fused_phase.argtypes = [_vp, _i, _vp, _vp, _i, _d, _d, _i, _i, _i, _i, _i64, _vp, _i, _i, _vp]
    lib.node_b200_fused_sums.argtypes = [_vp]
    lib.node_b200_fused_sums.restype = _vp
    lib.node_b200_fused_ctl.argtypes = [_vp]
    lib.node_b200_fused_ctl.restype = _vp
    lib.node_b200_vjp_workspace_bytes.argtypes = [_i, _i, _i, _i]
    lib.node_b200_vjp_workspace_bytes.restype = _i64
    lib.node_b200_odefunc_vjp.argtypes = [_vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]
    lib.node_b200_wgrad.argtypes = [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]
    lib.node_b200_vjp_buffer.argtypes = [_vp, _i, _i, _i, _i, _i]
    lib.node_b200_vjp_buffer.restype = _vp
    lib.node_b200_groupnorm_relu.argtypes = [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _f, _i, _vp]
    lib.node_b200_resconv_workspace_bytes.argtypes = [_i, _i, _i]
    lib.node_b200_resconv_workspace_bytes.restype = _i64
    lib.node_b200_resconv_prepare.argtypes = [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]
    lib.node_b200_resconv_forward.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]
    lib.node_b200_convs2_workspace_bytes.argtypes = [_i, _i, _i]
    lib.node_b200_convs2_workspace_bytes.restype = _i64
    lib.node_b200_convs2_prepare.argtypes = [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]
    lib.node_b200_convs2_forward.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]
    lib.node_b200_stem_gn_relu.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]
    lib.node_b200_head.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _f, _vp]
    lib.node_b200_feature_normalize.argtypes = [_vp, _vp, _vp, _i64, _i64, _i, _vp]
    lib.node_b200_retrieval_scores.argtypes = [_vp, _vp, _vp, _i64, _i64, _i, _vp]
    lib.node_b200_groupnorm_relu_backward.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _f, _i, _vp]
    lib.node_b200_absmax.argtypes = [_vp, _i64, _vp, _vp]
    lib.node_b200_plane_split.argtypes = [_vp, _vp, _i64, _i, _i, _i, _vp]
    lib.node_b200_plane_merge.argtypes = [_vp, _vp, _i64, _i, _i, _i, _vp]
    lib.node_b200_conv3x3_prepare.argtypes = [_vp, _i, _i, _i, _vp, _vp]
    lib.node_b200_conv3x3_forward.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]
    lib.node_b200_conv_wgrad_workspace_bytes.argtypes = [_i]
    lib.node_b200_conv_wgrad_workspace_bytes.restype = _i64
    lib.node_b200_conv_wgrad.argtypes = [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]
    lib.node_b200_stem_backward_workspace_bytes.argtypes = [_i]
    lib.node_b200_stem_backward_workspace_bytes.restype = _i64
    lib.node_b200_stem_backward.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]
    lib.node_b200_resconv_scal_offset.restype = _i64
    lib.node_b200_convs2_scal_offset.restype = _i64
    lib.node_b200_peer_alloc.argtypes = [_vp]
    lib.node_b200_peer_open.argtypes = [_i, _i, _vp]
    lib.node_b200_fold_reduce.argtypes = [_vp, _i, _vp, _i, _vp]
    lib.node_b200_conv3x3_forward_strided.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i64, _i64, _vp]
    lib.node_b200_wide_odefunc.argtypes = [_vp, _i64] + [_vp] * 15 + [_f, _i, _i, _i, _i, _vp]
    lib.node_b200_groupnorm_relu_ex.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _i64, _i, _i, _i, _f, _i, _vp]
    lib.node_b200_wide8_workspace_bytes.argtypes = [_i, _i, _i]
    lib.node_b200_wide8_workspace_bytes.restype = _i64
    lib.node_b200_wide8_operand_bytes.argtypes = [_i64, _i]
    lib.node_b200_wide8_operand_bytes.restype = _i64
    lib.node_b200_wide8_prepare.argtypes = [_vp, _i, _i, _i] + [_vp] * 7
    lib.node_b200_wide8_gn_operand.argtypes = [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _i, _vp]
    lib.node_b200_wide8_conv.argtypes = [_vp, _i, _vp, _vp, _i, _i, _vp]
    lib.node_b200_wide8_watchdog.argtypes = [_vp, _i, _vp]
    lib.node_b200_wide8_odefunc.argtypes = [_vp] * 14 + [_f, _i, _i, _vp]
    lib.node_b200_adjoint_step.argtypes = [_vp, _vp, _i64, _i, _vp, _vp, _i, _vp, _vp, _f, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]
    lib.node_b200_adjoint_solve.argtypes = [_vp, _vp, _i64, _vp, _vp, _i, _vp, _vp, _vp, _f, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]
    lib.node_b200_odefunc_vjp_split.argtypes = [_vp] * 5 + [_f] + [_vp] * 4 + [_i, _i, _i, _i, _vp, _vp, _vp]
    lib.node_b200_adjoint_solve_reset.argtypes = []
    lib.node_b200_groupnorm_backward_ex.argtypes = [_vp] * 8 + [_f, _vp, _vp, _vp, _i64, _i, _i, _i, _f, _i, _vp]
    lib.node_b200_batch_colsum.argtypes = [_vp, _vp, _i64, _i64, _vp]
    lib.node_b200_pow2_scale.argtypes = [_vp, _vp, _vp]
    lib.node_b200_wide_conv_blocks.argtypes = [_vp, _i64, _vp, _vp, _i, _i, _i, _i, _vp]
    lib.node_b200_wide_vjp.argtypes = [_vp, _vp, _f, _vp]
    lib.node_b200_lincomb.argtypes = [_i, _vp, _vp, _vp, _vp, _i, _i64, _vp]
    lib.node_b200_lincomb_scale.argtypes = [_i, _vp, _vp, _vp, _i, _i64, _vp]
    lib.node_b200_lincomb_dots.argtypes = [_i, _vp, _vp, _i, _i64, _vp, _vp, _vp]
    lib.node_b200_lincomb_scratch_doubles.restype = _i64
    lib.node_b200_lincomb_scratch_doubles.argtypes = []
    lib.node_b200_wide8_raw_operand.argtypes = [_vp, _i, _vp, _vp, _vp, _i, _i, _vp]


def lib():
    """Load (once) and return the shared library; raise loudly if it cannot serve."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'node_b200: native library %s is missing. There is no CPU or PyTorch fallback for the '
                'dopri5 hot path; build it with `python __graft_entry__.py build`.' % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        missing = [n for n in EXPORTS if not hasattr(handle, n)]
        if missing:
            raise RuntimeError('node_b200: %s lacks symbols %s - rebuild it' % (LIB_PATH, missing))
        _declare(handle)
        if handle.node_b200_abi_version() != ABI_VERSION:
            raise RuntimeError('node_b200: ABI mismatch - rebuild %s' % LIB_PATH)
        _lib = handle
    return _lib


def layout():
    global _layout
    if _layout is None:
        buf = (ctypes.c_int64 * 64)()
        n = lib().node_b200_ctl_layout(ctypes.cast(buf, _vp), 64)
        assert n == len(CTL_FIELDS), 'ctl layout mismatch between native.py and rk_kernels.cu'
        _layout = dict(zip(CTL_FIELDS, list(buf)[:n]))
    return _layout


def check(err, what):
    if err != 0:
        raise RuntimeError('node_b200: %s failed with CUDA error %d' % (what, err))


def on_device_of(argpos):
    """Decorator: run the wrapped entry point with the CUDA device of its `argpos`-th argument (a tensor, or a tuple whose
    first member is one) current. The kernels are launched on torch's CURRENT stream and the per-device kernel attributes
    are looked up through cudaGetDevice, so a tensor on cuda:1 must be served with cuda:1 current."""
    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*args, **kw):
            t = args[argpos]
            if isinstance(t, (tuple, list)) and t:
                t = t[0]
            if torch.is_tensor(t) and t.is_cuda and t.device.index != (_raw_device() if _raw_device is not None else torch.cuda.current_device()):
                with torch.cuda.device(t.device):
                    return fn(*args, **kw)
            return fn(*args, **kw)
        return wrapper
    return deco


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)
_raw_device = getattr(torch._C, '_cuda_getDevice', None)
_SLOW_STREAM = os.environ.get('NODE_B200_SLOW_STREAM', '0') == '1'       # A/B aid


def stream_ptr():
    """torch's CURRENT stream of the current device as a cudaStream_t. Every native call asks for it; the raw accessors cost ~1 us
    where `torch.cuda.current_stream()` builds a Stream object (18 us measured - 4 ms of a batch-128 training step)."""
    if _raw_stream is not None and _raw_device is not None and not _SLOW_STREAM:
        return ctypes.c_void_p(_raw_stream(_raw_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class device_guard(object):
    """`with device_guard(tensor.device):` - torch.cuda.device() only when the tensor's device is not the current one."""

    def __init__(self, device):
        cur = _raw_device() if _raw_device is not None else torch.cuda.current_device()
        self.ctx = torch.cuda.device(device) if (device.index is not None and device.index != cur) else None

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)
        return False


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def host_f64(values):
    """Host array of doubles that OWNS its memory: pass the returned object itself as the argument so
    it stays alive for the duration of the call (a bare .ctypes.data of a temporary would dangle)."""
    vals = [float(v) for v in np.asarray(values, dtype=np.float64).reshape(-1)]
    return (ctypes.c_double * max(len(vals), 1))(*vals)


def host_i64(values):
    vals = [int(v) for v in np.asarray(values, dtype=np.int64).reshape(-1)]
    return (ctypes.c_int64 * max(len(vals), 1))(*vals)


class CtlView(object):
    """Host snapshot of the device controller block (one D2H copy + sync)."""

    def __init__(self, ctl_dev_u8):
        self.raw = ctl_dev_u8.cpu().numpy()
        self.L = layout()

    def _get(self, name, dtype, count=1):
        off = self.L[name]
        v = np.frombuffer(self.raw, dtype=dtype, count=count, offset=off)
        return v if count > 1 else v[0].item()

    def i32(self, name):
        return self._get(name, np.int32)

    def f64(self, name, count=1):
        return self._get(name, np.float64, count)

    def trace(self):
        n = min(self.i32('n_attempt'), self.L['max_trace'])
        return dict(
            t=self.f64('tr_t', self.L['max_trace'])[:n].copy(),
            dt=self.f64('tr_dt', self.L['max_trace'])[:n].copy(),
            ratio=self.f64('tr_ratio', self.L['max_trace'])[:n].copy(),
            accepted=self._get('tr_acc', np.int32, self.L['max_trace'])[:n].astype(bool),
        )


def raise_for_status(status):
    """Same exception type and wording as the reference's asserts (dopri5.py:89,100,102; interp.py:58)."""
    if status == 0:
        return
    if status & ST_WATCHDOG:
        raise RuntimeError('node_b200: an in-kernel barrier wait timed out (status %d)' % status)
    if status & ST_DT_UNDERFLOW:
        raise AssertionError('underflow in dt')
    if status & ST_NONFINITE:
        raise AssertionError('non-finite values in state `y`')
    if status & ST_MAX_STEPS:
        raise AssertionError('max_num_steps exceeded')
    if status & ST_INTERP_RANGE:
        raise AssertionError('invalid interpolation, fails `t0 <= t <= t1`')
    raise AssertionError('solver failed with status %d' % status)
