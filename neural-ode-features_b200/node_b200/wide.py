"""Wide ODE-Net dynamics (n_filters = 128, 192, 256 - the paper's CIFAR setting, reference reproduce.sh:21, model.py:326-348)
on this repo's kernels instead of cuDNN.

The fused step engines are built for C = 64 (one accumulator tile = all output channels). For wider models the dynamics run
per evaluation as 64-channel blocks on the same tcgen05 machinery:

    a1 = relu(GN1(y))                                   node_b200_groupnorm_relu      (cells of C/32 channels)
    c1[co] = sum_ci conv3x3(a1[ci], W1[co, 1+ci])       node_b200_conv3x3_forward_strided, (C/64)^2 launches, `addend == out`
    a2 = relu(GN2(c1 + b1 + t * Tmap1))                 node_b200_groupnorm_relu_ex   (the time channel folded: model.py:320-323)
    c2[co] = sum_ci conv3x3(a2[ci], W2[co, 1+ci])
    k  = s * GN3(c2 + b2 + t * Tmap2)                   node_b200_groupnorm_relu_ex

(fp32 contract by fp16 operand splitting; the operand scale is found per super-tile inside the convolution kernel, so no
a-priori bound on the activations is needed). The Runge-Kutta stage combinations, error norm, controller and dense output
are the generic route's kernels - at C = 256 an evaluation is 16x the FLOPs of C = 64 for 4x the bytes, so leaving them
un-fused costs a few per cent. Served maps: 8x8 (CIFAR residual), 7x7 (MNIST residual), 15x15, 13x13.

8x8 maps with C = 128 / 256 (the CIFAR `residual` model) run on the wide8 engine instead (csrc/wide8_engine.cu): the GroupNorm pass
writes the activation once as the fp16 hi/lo operand image of the dense 8x8 tiling and each convolution is ONE TMA-fed tcgen05
implicit GEMM over all C channels (N = C accumulator columns, a weight tile read once per 4 images). `NODE_B200_WIDE8=0` keeps the
block path.
"""
import os
import weakref

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import native

_SERVED_HW = ((8, 8), (7, 7), (15, 15), (13, 13))


class WideAugmented(object):
    """The adjoint's augmented dynamics (adjoint.py:32-55) of a wide ODEfunc for the generic route (solver._GenericSolve): state
    (y, adj_y, adj_t, adj_params) -> (f, vjp_y, vjp_t, vjp_params), every convolution / GroupNorm (forward and backward) on this
    repo's kernels - no autograd graph, no cuDNN."""

    replicated_from = 2              # adj_t, adj_params are sums over the (possibly sharded) batch

    def __init__(self, func):
        self.func = func
        self.target = func.base if hasattr(func, 'base') else func
        self.dyn = WideDynamics.of(self.target)

    def eval_into(self, t_dev, src, dst, tsign):
        self.dyn.vjp_into(t_dev, src[0], src[1], dst, tsign)


def recognise(func):
    """The reference's ODEfunc (model.py:326-348) with C = 128, 192, 256, ... fp32 CUDA parameters; returns C or None."""
    if not isinstance(func, nn.Module) or (type(func).__name__ != 'ODEfunc' and not getattr(func, '_node_b200_fusable', False)):
        return None
    try:
        convs = [func.conv1._layer, func.conv2._layer]
        norms = [func.norm1, func.norm2, func.norm3]
    except AttributeError:
        return None
    if not isinstance(getattr(func, 'relu', None), nn.ReLU):
        return None
    C = convs[0].out_channels
    if C <= 64 or C % 64 != 0:
        return None
    for c in convs:
        if type(c) is not nn.Conv2d or (c.in_channels, c.out_channels, c.kernel_size, c.stride, c.padding, c.dilation, c.groups) != \
                (C + 1, C, (3, 3), (1, 1), (1, 1), (1, 1), 1) or c.bias is None or c.padding_mode != 'zeros':
            return None
        if c.weight.dtype != torch.float32 or not c.weight.is_cuda:
            return None
    for n in norms:
        if type(n) is not nn.GroupNorm or n.num_groups != 32 or n.num_channels != C or not n.affine or abs(n.eps - 1e-5) > 1e-12:
            return None
    return C


def serves(func, y0):
    C = recognise(func)
    if C is None or len(y0) != 1:
        return False
    y = y0[0]
    return (y.dtype == torch.float32 and y.is_cuda and y.dim() == 4 and y.shape[1] == C and tuple(y.shape[2:]) in _SERVED_HW
            and y.device == func.conv1._layer.weight.device and ((C // 32) * y.shape[2] * y.shape[3]) % 4 == 0)


class WideDynamics(object):
    """eval_into-style dynamics for the generic route (solver._GenericSolve): f(t, y) of a wide ODEfunc, block by block."""

    _cache = weakref.WeakKeyDictionary()

    def __init__(self, func):
        self.func = func
        self.C = recognise(func)
        self.nb = self.C // 64

    @classmethod
    def of(cls, func):
        inst = cls._cache.get(func)
        if inst is None:
            inst = cls._cache[func] = cls(func)
        return inst

    # ---- per parameter version: 64x64 weight blocks packed for the engine, time maps ------------------------------------
    def _prepare(self, H, W):
        convs = [self.func.conv1._layer, self.func.conv2._layer]
        key = (H, W) + tuple((c.weight.data_ptr(), c.weight._version) for c in convs)
        if getattr(self, '_key', None) == key:
            return
        lib = native.lib()
        dev = convs[0].weight.device
        nbytes = lib.node_b200_resconv_workspace_bytes(64, H, W)
        if not hasattr(self, '_ws') or self._ws.shape[-1] != nbytes or self._ws.device != dev:
            self._ws = torch.zeros((2, self.nb, self.nb, nbytes), dtype=torch.uint8, device=dev)
        self._tmap = []
        ones = torch.ones(1, 1, H, W, device=dev)
        for li, c in enumerate(convs):
            w = c.weight.detach()
            self._tmap.append(F.conv2d(ones, w[:, :1], padding=1)[0].contiguous())      # [C, H, W]: sum of the time-plane taps inside the map
            for co in range(self.nb):
                for ci in range(self.nb):
                    blk = w[64 * co:64 * co + 64, 1 + 64 * ci:1 + 64 * ci + 64].contiguous()
                    native.check(lib.node_b200_conv3x3_prepare(native.ptr(self._ws[li, co, ci]), 64, H, W, native.ptr(blk),
                                                               native.stream_ptr()), 'conv3x3_prepare')
        self._key = key

    # ---- wide8 engine: 8x8 maps, C = 128 / 256 ---------------------------------------------------------------------------
    def _wide8(self, H, W):
        return (H, W) == (8, 8) and self.C in (128, 256) and os.environ.get('NODE_B200_WIDE8', '1') != '0'

    def _prepare8(self):
        f = self.func
        ps = [f.conv1._layer.weight, f.conv2._layer.weight, f.norm1.weight, f.norm1.bias, f.norm2.weight, f.norm2.bias]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if getattr(self, '_key8', None) == key:
            return
        lib = native.lib()
        dev = ps[0].device
        nbytes = lib.node_b200_wide8_workspace_bytes(self.C, 8, 8)
        if not hasattr(self, '_ws8') or self._ws8.device != dev:
            self._ws8 = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        native.check(lib.node_b200_wide8_prepare(native.ptr(self._ws8), self.C, 8, 8, *[native.ptr(p.detach()) for p in ps],
                                                 native.stream_ptr()), 'wide8_prepare')
        self._key8 = key

    def _eval8(self, t_dev, y, out, tsign):
        N, C = int(y.shape[0]), self.C
        self._prepare8()
        f, lib = self.func, native.lib()
        nop = lib.node_b200_wide8_operand_bytes(N, C)
        if not hasattr(self, '_op8') or self._op8.numel() != nop or self._op8.device != y.device:
            self._op8 = torch.zeros(nop, dtype=torch.uint8, device=y.device)      # halo entries stay zero: the kernels never write them
            self._c8 = torch.empty_like(y)
        t32 = t_dev if t_dev.dtype == torch.float32 else t_dev.float()
        native.check(lib.node_b200_wide8_odefunc(
            native.ptr(self._ws8), native.ptr(y), native.ptr(out), native.ptr(self._op8), native.ptr(self._c8),
            native.ptr(f.norm1.weight), native.ptr(f.norm1.bias), native.ptr(f.norm2.weight), native.ptr(f.norm2.bias),
            native.ptr(f.norm3.weight), native.ptr(f.norm3.bias), native.ptr(f.conv1._layer.bias), native.ptr(f.conv2._layer.bias),
            native.ptr(t32), float(tsign), N, C, native.stream_ptr()), 'wide8_odefunc')

    def watchdog(self):
        """1 if a bounded barrier wait of the wide8 convolution kernel expired since the last prepare (synchronises)."""
        return int(native.lib().node_b200_wide8_watchdog(native.ptr(self._ws8), self.C, native.stream_ptr())) if hasattr(self, '_ws8') else 0

    def eval_into(self, t_dev, src, dst, tsign):
        """dst[0] = tsign * f(tsign * t, src[0]) (misc.py:184-187 for reversed time)."""
        y, out = src[0], dst[0]
        N, C, H, W = (int(v) for v in y.shape)
        if self._wide8(H, W):
            self._eval8(t_dev, y, out, tsign)
            if hasattr(self.func, 'nfe'):
                self.func.nfe += 1
            return
        self._prepare(H, W)
        f, lib, sp = self.func, native.lib(), native.stream_ptr
        if not hasattr(self, '_tmp') or self._tmp.shape[1:] != y.shape or self._tmp.device != y.device:
            self._tmp = torch.empty((2,) + tuple(y.shape), dtype=y.dtype, device=y.device)
        a, c = self._tmp[0], self._tmp[1]
        t32 = t_dev if t_dev.dtype == torch.float32 else t_dev.float()
        native.check(lib.node_b200_wide_odefunc(
            native.ptr(self._ws), int(self._ws.shape[-1]), native.ptr(y), native.ptr(out), native.ptr(a), native.ptr(c),
            native.ptr(f.norm1.weight), native.ptr(f.norm1.bias), native.ptr(f.norm2.weight), native.ptr(f.norm2.bias),
            native.ptr(f.norm3.weight), native.ptr(f.norm3.bias), native.ptr(f.conv1._layer.bias), native.ptr(self._tmap[0]),
            native.ptr(f.conv2._layer.bias), native.ptr(self._tmap[1]), native.ptr(t32), float(tsign), N, C, H, W, sp()), 'wide_odefunc')
        if hasattr(f, 'nfe'):
            f.nfe += 1                                       # model.py:340 counts every evaluation

    # ---- the adjoint's augmented dynamics (adjoint.py:32-55) -------------------------------------------------------------
    def _prepare_vjp(self, H, W):
        """Per parameter version: the data-gradient weight blocks (conv3x3 with W^T and flipped taps), validity matrix of the
        folded time channel."""
        self._prepare(H, W)
        convs = [self.func.conv1._layer, self.func.conv2._layer]
        key = (H, W) + tuple((c.weight.data_ptr(), c.weight._version) for c in convs)
        if getattr(self, '_key_vjp', None) == key:
            return
        lib = native.lib()
        dev = convs[0].weight.device
        nbytes = lib.node_b200_resconv_workspace_bytes(64, H, W)
        if not hasattr(self, '_wsd') or self._wsd.shape[-1] != nbytes or self._wsd.device != dev:
            self._wsd = torch.zeros((2, self.nb, self.nb, nbytes), dtype=torch.uint8, device=dev)
        for li, c in enumerate(convs):
            wd = c.weight.detach()[:, 1:].flip(2, 3).transpose(0, 1)      # [ci, co, 3, 3]: dL/dx = conv3x3(dL/dc, wd)
            for o in range(self.nb):
                for i in range(self.nb):
                    blk = wd[64 * o:64 * o + 64, 64 * i:64 * i + 64].contiguous()
                    native.check(lib.node_b200_conv3x3_prepare(native.ptr(self._wsd[li, o, i]), 64, H, W, native.ptr(blk),
                                                               native.stream_ptr()), 'conv3x3_prepare')
        ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing='ij')
        V = torch.zeros(H * W, 9, device=dev)
        for ky in range(3):
            for kx in range(3):
                ok = (ys + ky - 1 >= 0) & (ys + ky - 1 < H) & (xs + kx - 1 >= 0) & (xs + kx - 1 < W)
                V[:, ky * 3 + kx] = ok.reshape(-1).float()
        self._valid = V                                                    # Tmap[c] = W[c, 0].view(9) @ V^T
        self._key_vjp = key

    def _prepare8_dgrad(self):
        """wide8 workspace of the data gradients: the same engine on W^T with flipped taps (the time plane has no data gradient)."""
        f = self.func
        ps = [f.conv1._layer.weight, f.conv2._layer.weight]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if getattr(self, '_key8d', None) == key:
            return
        lib = native.lib()
        dev = ps[0].device
        nbytes = lib.node_b200_wide8_workspace_bytes(self.C, 8, 8)
        if not hasattr(self, '_ws8d') or self._ws8d.device != dev:
            self._ws8d = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        wd = []
        for p in ps:
            full = torch.zeros_like(p)
            full[:, 1:] = p.detach()[:, 1:].flip(2, 3).transpose(0, 1)
            wd.append(full)
        native.check(lib.node_b200_wide8_prepare(native.ptr(self._ws8d), self.C, 8, 8, native.ptr(wd[0]), native.ptr(wd[1]),
                                                 native.ptr(f.norm1.weight), native.ptr(f.norm1.bias), native.ptr(f.norm2.weight),
                                                 native.ptr(f.norm2.bias), native.stream_ptr()), 'wide8_prepare (data gradient)')
        self._key8d = key

    def vjp_into(self, t_dev, y, adj_y, dst, tsign):
        """dst = tsign * (f, vjp_y, vjp_t, vjp_params)(tsign * t) of ODEfunc with cotangent -adj_y (adjoint.py:40-49; a reversed span
        negates the whole augmented system and its time argument, misc.py:184-187), parameters in `func.parameters()` order
        (misc.py:5-7) - one C call (node_b200_wide_vjp, csrc/wide_vjp.cu)."""
        f, lib = self.func, native.lib()
        N, C, H, W = (int(v) for v in y.shape)
        w8 = self._wide8(H, W)
        if w8:
            self._prepare8()
            self._prepare8_dgrad()
            if getattr(self, '_tmap8_key', None) != self._key8:          # time maps for the GroupNorm / bias kernels (fp32 [C, 8, 8])
                ones = torch.ones(1, 1, H, W, device=y.device)
                self._tmap = [F.conv2d(ones, c.weight.detach()[:, :1], padding=1)[0].contiguous() for c in (f.conv1._layer, f.conv2._layer)]
                self._tmap8_key = self._key8
        else:
            self._prepare_vjp(H, W)
        v = getattr(self, '_v', None)
        if v is None or v['a1'].shape != y.shape or v['a1'].device != y.device:
            e = lambda: torch.empty_like(y)
            dev = y.device
            v = self._v = dict(a1=e(), c1=e(), a2=e(), c2=e(), gc2=e(), gr=e(), gc1=e(),
                               ab=torch.empty((self.nb, N, 64, H, W), device=dev), gb=torch.empty((self.nb, N, 64, H, W), device=dev),
                               part=torch.empty(2 * N * C, device=dev), S=torch.empty(C * H * W, device=dev),
                               bits=torch.zeros(2, dtype=torch.int32, device=dev), scale=torch.ones(1, device=dev),
                               wg=torch.empty(max(lib.node_b200_conv_wgrad_workspace_bytes(k) for k in range(1, 7)), dtype=torch.uint8, device=dev),
                               dw=torch.empty(self.nb * self.nb * 64 * 64 * 9, device=dev),
                               vt=torch.zeros(2 * C, dtype=torch.float64, device=dev))
            if w8:
                v['op'] = torch.zeros(lib.node_b200_wide8_operand_bytes(N, C), dtype=torch.uint8, device=dev)   # halo entries stay zero
        t32 = t_dev if t_dev.dtype == torch.float32 else t_dev.float()
        n1, n2, n3, cv1, cv2 = f.norm1, f.norm2, f.norm3, f.conv1._layer, f.conv2._layer
        tensors = [y, adj_y, t32, dst[0], dst[1], dst[2], dst[3], n1.weight, n1.bias, n2.weight, n2.bias, n3.weight, n3.bias,
                   cv1.bias, cv2.bias, self._tmap[0], self._tmap[1], None if w8 else self._ws, None if w8 else self._wsd,
                   self._ws8 if w8 else None, self._ws8d if w8 else None, v.get('op'), v['a1'], v['c1'], v['a2'], v['c2'], v['gc2'],
                   v['gr'], v['gc1'], v['ab'], v['gb'], v['part'], v['S'], v['bits'], v['scale'], v['wg'], v['dw'], v['vt']]
        for x in (y, adj_y, dst[0], dst[1], dst[3]):
            assert x.is_contiguous()
        ptrs = (native._vp * len(tensors))(*[x.data_ptr() if x is not None else 0 for x in tensors])
        dims = native.host_i64([N, C, H, W, 0 if w8 else int(self._ws.shape[-1]), 1 if w8 else 0])
        native.check(lib.node_b200_wide_vjp(ptrs, dims, float(tsign), native.stream_ptr()), 'wide_vjp')
        if hasattr(f, 'nfe'):
            f.nfe += 1                                       # model.py:340 counts every evaluation

    def __call__(self, t, y):
        """f(t, y) as a plain function (tests): y a tensor, t a python float or 0-d tensor."""
        t_dev = t if torch.is_tensor(t) else torch.tensor(float(t), dtype=torch.float32, device=y.device)
        out = torch.empty_like(y)
        self.eval_into(t_dev.to(y.device), (y.contiguous(),), (out,), 1.0)
        return out
