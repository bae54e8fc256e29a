"""Training path of the callers: the residual downsampler and the GroupNorm -> ReLU pairs under autograd
(reference model.py:119-178, 231-250 inside train.py:40-58) on this repo's kernels, forward AND backward.

With gradients enabled the reference's modules run cuDNN fp32 convolutions (SIMT / FFT algorithms) and ATen GroupNorm; here
every fused caller kernel of caller_ops.py is wrapped in an autograd Function whose backward is native too:

  * GroupNorm -> ReLU:          csrc/caller_bwd.cu k_gn_relu_bwd (2 reads + 1 write, statistics recomputed).
  * ResBlock tail  r = conv2(relu(norm2(c))) + shortcut:
        data gradient  = the same tcgen05 implicit GEMM on the flipped, transposed kernel (node_b200_conv3x3_forward),
        weight gradient = the adjoint's weight-gradient GEMM (node_b200_conv_wgrad) on (relu(norm2(c)), grad),
        then the GroupNorm -> ReLU backward.
  * ResBlock head  (conv1 3x3 stride 2, 1x1 stride 2 shortcut): on the four parity planes of the input a stride-2 tap is a
    stride-1 tap, so both gradients are the stride-1 engines applied per plane (plane_split / plane_merge).
  * Stem  relu(norm(conv0(x))): node_b200_stem_backward recomputes the convolution per image and never materialises
    its output or its gradient. The input x receives no gradient on this path (callers that need dL/dx - the PGD
    attack - keep the modules' own ops).

`NODE_B200_CALLER_GRAD=0` switches all of it off (the modules' own PyTorch ops, as in the reference).
"""
import ctypes
import os

import torch

from . import native

launches = 0


def enabled():
    return os.environ.get('NODE_B200_CALLER_GRAD', '1') != '0' and os.environ.get('NODE_B200_CALLERS', '1') != '0'


def _count(n=1):
    global launches
    launches += n


def _scalar_ptr(buf, offset):
    return ctypes.c_void_p(buf.data_ptr() + offset)


def _ptr_array(tensors_or_ptrs):
    vals = [(t.data_ptr() if torch.is_tensor(t) else (t.value if isinstance(t, ctypes.c_void_p) else int(t))) for t in tensors_or_ptrs]
    return (ctypes.c_void_p * len(vals))(*vals)


# ---- building blocks -------------------------------------------------------------------------------------------------------

def gn_relu_forward(x, w, b, groups, eps, relu=True):
    y = torch.empty_like(x)
    N, C = int(x.shape[0]), int(x.shape[1])
    native.check(native.lib().node_b200_groupnorm_relu(native.ptr(x), native.ptr(y), native.ptr(w), native.ptr(b), N, C, groups,
                                                       x.numel() // (N * C), float(eps), 1 if relu else 0, native.stream_ptr()),
                 'groupnorm_relu')
    _count()
    return y


def gn_relu_backward(x, gy, w, b, groups, eps, relu=True):
    """(dL/dx, dL/dgamma, dL/dbeta) of y = relu?(GroupNorm(x))."""
    N, C = int(x.shape[0]), int(x.shape[1])
    gx = torch.empty_like(x)
    part = torch.empty(2 * N * C, dtype=torch.float32, device=x.device)
    gw = torch.empty(C, dtype=torch.float32, device=x.device)
    gb = torch.empty(C, dtype=torch.float32, device=x.device)
    native.check(native.lib().node_b200_groupnorm_relu_backward(
        native.ptr(x), native.ptr(gy), native.ptr(gx), native.ptr(w), native.ptr(b), native.ptr(part), native.ptr(gw), native.ptr(gb),
        N, C, groups, x.numel() // (N * C), float(eps), 1 if relu else 0, native.stream_ptr()), 'groupnorm_relu_backward')
    _count(3)
    return gx, gw, gb


def absmax_bits(t):
    out = torch.empty(1, dtype=torch.int32, device=t.device)
    native.check(native.lib().node_b200_absmax(native.ptr(t), t.numel(), native.ptr(out), native.stream_ptr()), 'absmax')
    _count()
    return out


def plane_split(a):
    N, C, HI, WI = (int(v) for v in a.shape)
    HO, WO = (HI - 1) // 2 + 1, (WI - 1) // 2 + 1
    planes = torch.empty((4, N, C, HO, WO), dtype=a.dtype, device=a.device)
    native.check(native.lib().node_b200_plane_split(native.ptr(a), native.ptr(planes), N, C, HI, WI, native.stream_ptr()), 'plane_split')
    _count()
    return planes


def plane_merge(gplanes, HI, WI):
    _, N, C, _, _ = (int(v) for v in gplanes.shape)
    ga = torch.empty((N, C, HI, WI), dtype=gplanes.dtype, device=gplanes.device)
    native.check(native.lib().node_b200_plane_merge(native.ptr(gplanes), native.ptr(ga), N, C, HI, WI, native.stream_ptr()), 'plane_merge')
    _count()
    return ga


_raw_ws = {}


def conv3x3_raw(x, weight, addend=None, slot=0, out=None):
    """conv2d(x, weight, stride 1, padding 1) (+ addend) for a signed x [N,64,H,W] on the tcgen05 engine. `slot` separates the
    prepared-weight buffers of the calls of one backward pass (they are re-packed on every call: the weights change every step)."""
    N, C, H, W = (int(v) for v in x.shape)
    lib = native.lib()
    key = (str(x.device), H, W, slot)
    buf = _raw_ws.get(key)
    if buf is None:
        buf = _raw_ws[key] = torch.zeros(lib.node_b200_resconv_workspace_bytes(C, H, W), dtype=torch.uint8, device=x.device)
    weight = weight.contiguous()
    native.check(lib.node_b200_conv3x3_prepare(native.ptr(buf), C, H, W, native.ptr(weight), native.stream_ptr()), 'conv3x3_prepare')
    out = torch.empty_like(x) if out is None else out
    native.check(lib.node_b200_conv3x3_forward(native.ptr(buf), native.ptr(x), native.ptr(addend), native.ptr(out), N, C, H, W,
                                               native.stream_ptr()), 'conv3x3_forward')
    _count(3)
    return out


_wgrad_ws = {}


def conv_wgrad(inputs, grads, input_scale_ptrs):
    """dW[p] [64,64,3,3] of 3x3 stride-1 padding-1 convolutions for pairs (inputs[p], grads[p]) of [N,64,H,W] tensors; the
    inputs are non-negative activations whose power-of-two operand scale is the device scalar input_scale_ptrs[p]."""
    n = len(inputs)
    N, C, H, W = (int(v) for v in inputs[0].shape)
    lib = native.lib()
    dev = inputs[0].device
    key = (str(dev), n)
    ws = _wgrad_ws.get(key)
    if ws is None:
        ws = _wgrad_ws[key] = torch.empty(lib.node_b200_conv_wgrad_workspace_bytes(n), dtype=torch.uint8, device=dev)
    bits = {}
    for g in grads:                                   # one |max| reduction per distinct gradient tensor
        if g.data_ptr() not in bits:
            bits[g.data_ptr()] = absmax_bits(g)
    dw = torch.empty((n, 64, 64, 3, 3), dtype=torch.float32, device=dev)
    native.check(lib.node_b200_conv_wgrad(native.ptr(ws), n, _ptr_array(inputs), _ptr_array(grads), _ptr_array(input_scale_ptrs),
                                          _ptr_array([bits[g.data_ptr()] for g in grads]), native.ptr(dw), N, C, H, W,
                                          native.stream_ptr()), 'conv_wgrad')
    _count(2)
    return dw


def _flip_t(w):
    """The kernel of the data gradient of a stride-1 padding-1 3x3 convolution with weight w [co,ci,3,3]."""
    return w.flip(2, 3).transpose(0, 1).contiguous()


# stride-2 tap (dy, dx) of conv1 = tap (ky, kx) of a stride-1 convolution on parity plane (pr, pc): rows 2i+dy-1 -> dy = 0: plane 1 at
# row i-1 (ky 0); dy = 1: plane 0 at row i (ky 1); dy = 2: plane 1 at row i (ky 1); the same for columns.
_PLANE_OF = {0: (1, 0), 1: (0, 1), 2: (1, 1)}        # d -> (parity, k)


def _plane_kernels(wc):
    """K[p] [co,ci,3,3] with conv1(a) = sum_p conv_s1(plane_p(a), K[p]), p = 2*pr + pc."""
    ks = wc.new_zeros((4,) + tuple(wc.shape))
    for dy in range(3):
        pr, ky = _PLANE_OF[dy]
        for dx in range(3):
            pc, kx = _PLANE_OF[dx]
            ks[2 * pr + pc, :, :, ky, kx] = wc[:, :, dy, dx]
    return ks


def _from_plane_grads(dks):
    """Inverse of _plane_kernels for the weight gradient: dWc[:, :, dy, dx] = dK[plane(dy, dx)][:, :, ky, kx]."""
    dwc = dks.new_zeros((64, 64, 3, 3))
    for dy in range(3):
        pr, ky = _PLANE_OF[dy]
        for dx in range(3):
            pc, kx = _PLANE_OF[dx]
            dwc[:, :, dy, dx] = dks[2 * pr + pc, :, :, ky, kx]
    return dwc


# ---- autograd Functions -----------------------------------------------------------------------------------------------------

class GnRelu(torch.autograd.Function):
    """y = relu?(GroupNorm(groups, C)(x)), C = 2 * groups."""

    @staticmethod
    def forward(ctx, x, w, b, groups, eps, relu):
        x = x.contiguous()
        ctx.save_for_backward(x, w, b)
        ctx.cfg = (groups, eps, relu)
        return gn_relu_forward(x, w, b, groups, eps, relu)

    @staticmethod
    def backward(ctx, gy):
        x, w, b = ctx.saved_tensors
        groups, eps, relu = ctx.cfg
        gx, gw, gb = gn_relu_backward(x, gy.contiguous(), w, b, groups, eps, relu)
        return gx, gw, gb, None, None, None


class Stem(torch.autograd.Function):
    """o = relu(GroupNorm(32, 64)(conv0(x))); no gradient for x."""

    @staticmethod
    def forward(ctx, x, cw, cb, gw, gb, eps):
        x = x.contiguous()
        N, CIN, HIN, WIN = (int(v) for v in x.shape)
        out = torch.empty((N, 64, HIN - 2, WIN - 2), dtype=x.dtype, device=x.device)
        native.check(native.lib().node_b200_stem_gn_relu(native.ptr(x), native.ptr(cw), native.ptr(cb), native.ptr(gw), native.ptr(gb),
                                                         native.ptr(out), N, CIN, HIN, WIN, float(eps), native.stream_ptr()), 'stem_gn_relu')
        _count()
        ctx.save_for_backward(x, cw, cb, gw, gb)
        ctx.eps = eps
        return out

    @staticmethod
    def backward(ctx, go):
        x, cw, cb, gw, gb = ctx.saved_tensors
        N, CIN, HIN, WIN = (int(v) for v in x.shape)
        lib = native.lib()
        ws = torch.empty(lib.node_b200_stem_backward_workspace_bytes(CIN), dtype=torch.uint8, device=x.device)
        grads = torch.empty(64 * CIN * 9 + 3 * 64, dtype=torch.float32, device=x.device)
        native.check(lib.node_b200_stem_backward(native.ptr(x), native.ptr(cw), native.ptr(cb), native.ptr(gw), native.ptr(gb),
                                                 native.ptr(go.contiguous()), native.ptr(ws), native.ptr(grads), N, CIN, HIN, WIN,
                                                 float(ctx.eps), native.stream_ptr()), 'stem_backward')
        _count(2)
        k = 64 * CIN * 9
        return None, grads[:k].view_as(cw), grads[k:k + 64], grads[k + 64:k + 128], grads[k + 128:k + 192], None


class ResTail(torch.autograd.Function):
    """r = conv2(relu(norm2(c))) + shortcut (model.py:175-178)."""

    @staticmethod
    def forward(ctx, c, shortcut, gw, gb, cw, eps, ws_buf):
        c, shortcut = c.contiguous(), shortcut.contiguous()
        N, C, H, W = (int(v) for v in c.shape)
        out = torch.empty_like(c)
        native.check(native.lib().node_b200_resconv_forward(native.ptr(ws_buf), native.ptr(c), native.ptr(shortcut), native.ptr(out),
                                                            native.ptr(gw), native.ptr(gb), None, None, N, C, H, W, float(eps),
                                                            native.stream_ptr()), 'resconv_forward')
        _count()
        ctx.save_for_backward(c, gw, gb, cw)
        ctx.eps, ctx.ws_buf = eps, ws_buf
        return out

    @staticmethod
    def backward(ctx, gr):
        c, gw, gb, cw = ctx.saved_tensors
        gr = gr.contiguous()
        o = gn_relu_forward(c, gw, gb, 32, ctx.eps, True)                                  # recomputed: 1 read + 1 write
        scale = _scalar_ptr(ctx.ws_buf, native.lib().node_b200_resconv_scal_offset())     # the forward's activation scale
        dw = conv_wgrad([o], [gr], [scale])[0]
        go = conv3x3_raw(gr, _flip_t(cw), slot=0)
        gc, dgw, dgb = gn_relu_backward(c, go, gw, gb, 32, ctx.eps, True)
        return gc, gr, dgw, dgb, dw, None, None


class ResHead(torch.autograd.Function):
    """(conv1(a), downsample(a)) for the strided ResBlock head (model.py:170-174): conv1 3x3 stride 2 padding 1, 1x1 stride 2."""

    @staticmethod
    def forward(ctx, a, wc, wd, ws_buf):
        a = a.contiguous()
        N, C, HI, WI = (int(v) for v in a.shape)
        HO, WO = (HI - 1) // 2 + 1, (WI - 1) // 2 + 1
        c = torch.empty((N, C, HO, WO), dtype=a.dtype, device=a.device)
        sc = torch.empty_like(c)
        native.check(native.lib().node_b200_convs2_forward(native.ptr(ws_buf), native.ptr(a), native.ptr(c), native.ptr(sc), N, C, HI, WI,
                                                           native.stream_ptr()), 'convs2_forward')
        _count()
        ctx.save_for_backward(a, wc, wd)
        ctx.ws_buf = ws_buf
        return c, sc

    @staticmethod
    def backward(ctx, gc, gsc):
        a, wc, wd = ctx.saved_tensors
        N, C, HI, WI = (int(v) for v in a.shape)
        gc, gsc = gc.contiguous(), gsc.contiguous()
        planes = plane_split(a)
        scale = _scalar_ptr(ctx.ws_buf, native.lib().node_b200_convs2_scal_offset())
        dks = conv_wgrad([planes[0], planes[1], planes[2], planes[3], planes[0]], [gc, gc, gc, gc, gsc], [scale] * 5)
        dwc = _from_plane_grads(dks[:4])
        dwd = dks[4][:, :, 1, 1].reshape(wd.shape)
        ks = _plane_kernels(wc)
        ksc = wc.new_zeros(wc.shape)
        ksc[:, :, 1, 1] = wd.reshape(64, 64)
        gplanes = torch.empty_like(planes)
        t = conv3x3_raw(gsc, _flip_t(ksc), slot=1)
        conv3x3_raw(gc, _flip_t(ks[0]), addend=t, slot=2, out=gplanes[0])
        for p in range(1, 4):
            conv3x3_raw(gc, _flip_t(ks[p]), slot=2 + p, out=gplanes[p])
        return plane_merge(gplanes, HI, WI), dwc, dwd, None
