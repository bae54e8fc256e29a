"""Callers of the hot path: GroupNorm -> ReLU of the downsamplers and the classifier head
(reference model.py:119-178, 231-250, 268-271) as one memory-bound CUDA pass (csrc/caller_ops.cu).

Used by `node_b200.models` for inference (no autograd graph); with gradients enabled the modules run their own
PyTorch ops, exactly as in the reference. The module tree - and so the state_dict keys - is unchanged."""
import os
import weakref

import torch
import torch.nn as nn

from . import caller_grad, native

launches = 0        # kernels launched by this module (bench.py's gpu_launches)


def enabled():
    """NODE_B200_CALLERS=0: every caller kernel off - the modules run the reference's own PyTorch ops (bench.py's incumbent arm)."""
    return os.environ.get('NODE_B200_CALLERS', '1') != '0'


def _fusable(norm, x):
    if not enabled():
        return False
    if not isinstance(norm, nn.GroupNorm) or norm.weight is None or norm.bias is None:
        return False
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() >= 3 and norm.weight.is_cuda):
        return False
    cell = (x.shape[1] // norm.num_groups) * (x.numel() // (x.shape[0] * x.shape[1]))
    if _needs_grad(x, norm.weight, norm.bias):       # training: the native backward serves GroupNorm(g, 2g) cells of <= 2048 floats
        if not (caller_grad.enabled() and x.shape[1] == 2 * norm.num_groups and cell <= 2048):
            return False
    return x.shape[1] == norm.num_channels and 0 < cell <= 4096 and x.shape[0] > 0


def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


@native.on_device_of(1)
def group_norm_relu(norm, x, relu=True):
    """relu(norm(x)) (or norm(x)) for an nn.GroupNorm `norm`; one CUDA pass when no gradient is needed."""
    if not _fusable(norm, x):
        y = norm(x)
        return torch.relu(y) if relu else y
    if _needs_grad(x, norm.weight, norm.bias):
        return caller_grad.GnRelu.apply(x, norm.weight, norm.bias, int(norm.num_groups), float(norm.eps), bool(relu))
    x = x.contiguous()
    y = torch.empty_like(x)
    N, C = int(x.shape[0]), int(x.shape[1])
    err = native.lib().node_b200_groupnorm_relu(native.ptr(x), native.ptr(y), native.ptr(norm.weight), native.ptr(norm.bias),
                                               N, C, int(norm.num_groups), x.numel() // (N * C), float(norm.eps),
                                               1 if relu else 0, native.stream_ptr())
    native.check(err, 'groupnorm_relu')
    global launches
    launches += 1
    return y


def run_sequential(seq, x):
    """nn.Sequential.forward with every (GroupNorm, ReLU) pair fused."""
    mods = list(seq.children())
    i = 0
    while i < len(mods):
        m = mods[i]
        if (isinstance(m, nn.Conv2d) and i + 2 < len(mods) and isinstance(mods[i + 1], nn.GroupNorm)
                and isinstance(mods[i + 2], nn.ReLU) and _stem_ok(m, mods[i + 1], x)):
            x = stem_gn_relu(m, mods[i + 1], x)          # stem convolution + GroupNorm + ReLU in one pass
            i += 3
            continue
        if isinstance(m, nn.GroupNorm):
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if isinstance(nxt, nn.ReLU):
                x = group_norm_relu(m, x, relu=True)
                i += 2
                continue
            x = group_norm_relu(m, x, relu=False)
        elif type(m) is nn.Sequential:
            x = run_sequential(m, x)
        else:
            x = m(x)
        i += 1
    return x


# ---- fused ResBlock tail: conv2(relu(norm2(x))) + shortcut (csrc/resconv_engine.cuh) ---------------------------------

_resconv_ws = {}


def _resconv_ok(norm, conv, x, shortcut):
    if not enabled():
        return False
    if not (isinstance(norm, nn.GroupNorm) and isinstance(conv, nn.Conv2d)) or norm.weight is None:
        return False
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and shortcut.shape == x.shape and shortcut.dtype == x.dtype):
        return False
    if _needs_grad(x, shortcut, conv.weight, norm.weight, norm.bias) and not (caller_grad.enabled() and x.shape[2:] in ((15, 15), (8, 8), (13, 13), (7, 7))):
        return False
    if (conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, conv.dilation, conv.groups) != \
            (64, 64, (3, 3), (1, 1), (1, 1), (1, 1), 1) or conv.bias is not None or conv.padding_mode != 'zeros':
        return False
    if norm.num_groups != 32 or norm.num_channels != 64 or x.shape[1] != 64:
        return False
    return native.lib().node_b200_resconv_workspace_bytes(64, int(x.shape[2]), int(x.shape[3])) > 0


@native.on_device_of(2)
def res_conv(norm, conv, x, shortcut, next_norm=None):
    """conv(relu(norm(x))) + shortcut for the ResBlock tail; one tcgen05 kernel when the shape is served and no gradient
    is needed, the modules' own ops otherwise. With `next_norm` (the following block's norm1) the result is
    relu(next_norm(.)) of that - fused into the same kernel when next_norm is a GroupNorm(32, 64) with the same eps."""
    training = _needs_grad(x, shortcut, conv.weight, norm.weight, norm.bias if isinstance(norm, nn.GroupNorm) else None,
                           next_norm.weight if isinstance(next_norm, nn.GroupNorm) else None)
    fuse_next = (next_norm is not None and isinstance(next_norm, nn.GroupNorm) and next_norm.weight is not None
                 and next_norm.num_groups == 32 and next_norm.num_channels == 64 and isinstance(norm, nn.GroupNorm)
                 and next_norm.eps == norm.eps and not training)       # training: the block output itself is needed by the backward
    if not _resconv_ok(norm, conv, x, shortcut):
        out = conv(torch.relu(norm(x))) + shortcut
        return out if next_norm is None else group_norm_relu(next_norm, out)
    x, shortcut = x.contiguous(), shortcut.contiguous()
    N, C, H, W = (int(v) for v in x.shape)
    lib = native.lib()
    key = (id(conv), id(norm), str(x.device), H, W)
    ent = _resconv_ws.get(key)
    ver = (conv.weight.data_ptr(), conv.weight._version, norm.weight.data_ptr(), norm.weight._version, norm.bias._version)
    live = (conv.weight, norm.weight, norm.bias)
    if ent is not None and not all(r() is p for r, p in zip(ent[2], live)):      # id() reused by a new module
        ent = None
    if ent is None or ent[1] != ver:
        if len(_resconv_ws) > 32:
            _resconv_ws.clear()
        buf = ent[0] if ent is not None else torch.zeros(lib.node_b200_resconv_workspace_bytes(C, H, W), dtype=torch.uint8, device=x.device)
        native.check(lib.node_b200_resconv_prepare(native.ptr(buf), C, H, W, native.ptr(conv.weight), native.ptr(norm.weight),
                                                   native.ptr(norm.bias), native.stream_ptr()), 'resconv_prepare')
        ent = _resconv_ws[key] = (buf, ver, [weakref.ref(p) for p in live])
    if training:
        out = caller_grad.ResTail.apply(x, shortcut, norm.weight, norm.bias, conv.weight, float(norm.eps), ent[0])
        return out if next_norm is None else group_norm_relu(next_norm, out)
    out = torch.empty_like(x)
    nw = native.ptr(next_norm.weight) if fuse_next else None
    nb = native.ptr(next_norm.bias) if fuse_next else None
    native.check(lib.node_b200_resconv_forward(native.ptr(ent[0]), native.ptr(x), native.ptr(shortcut), native.ptr(out),
                                               native.ptr(norm.weight), native.ptr(norm.bias), nw, nb, N, C, H, W, float(norm.eps),
                                               native.stream_ptr()), 'resconv_forward')
    global launches
    launches += 1
    if next_norm is not None and not fuse_next:
        return group_norm_relu(next_norm, out)
    return out


# ---- fused strided ResBlock head: conv1(a) [3x3 stride 2] and downsample(a) [1x1 stride 2] (csrc/convs2_engine.cuh) -------

_convs2_ws = {}


def _is_conv(m, k, stride, pad):
    return (isinstance(m, nn.Conv2d) and (m.in_channels, m.out_channels, m.kernel_size, m.stride, m.padding, m.dilation, m.groups)
            == (64, 64, (k, k), (stride, stride), (pad, pad), (1, 1), 1) and m.bias is None and m.padding_mode == 'zeros')


def _convs2_ok(norm, conv, down, a):
    if not enabled():
        return False
    if not (isinstance(norm, nn.GroupNorm) and norm.weight is not None and _is_conv(conv, 3, 2, 1) and _is_conv(down, 1, 2, 0)):
        return False
    if not (a.is_cuda and a.dtype == torch.float32 and a.dim() == 4 and a.shape[1] == 64 and norm.num_groups == 32):
        return False
    if _needs_grad(a, conv.weight, down.weight) and not (caller_grad.enabled() and a.shape[2:] in ((30, 30), (15, 15), (26, 26), (13, 13))):
        return False
    return native.lib().node_b200_convs2_workspace_bytes(64, int(a.shape[2]), int(a.shape[3])) > 0


@native.on_device_of(3)
def res_head(norm, conv, down, a):
    """(conv(a), down(a)) for a = relu(norm(x)) already computed: one tcgen05 kernel when the shape is served and no gradient
    is needed, the modules' own ops otherwise. `norm` only supplies the bound of |a| for the fp16 operand split."""
    if not _convs2_ok(norm, conv, down, a):
        return conv(a), down(a)
    a = a.contiguous()
    N, C, HI, WI = (int(v) for v in a.shape)
    HO, WO = (HI - 1) // 2 + 1, (WI - 1) // 2 + 1
    lib = native.lib()
    key = (id(conv), id(down), id(norm), str(a.device), HI, WI)
    ent = _convs2_ws.get(key)
    ver = (conv.weight.data_ptr(), conv.weight._version, down.weight.data_ptr(), down.weight._version, norm.weight._version,
           norm.bias._version)
    live = (conv.weight, down.weight, norm.weight, norm.bias)
    if ent is not None and not all(r() is p for r, p in zip(ent[2], live)):      # id() reused by a new module
        ent = None
    if ent is None or ent[1] != ver:
        if len(_convs2_ws) > 32:
            _convs2_ws.clear()
        buf = ent[0] if ent is not None else torch.zeros(lib.node_b200_convs2_workspace_bytes(C, HI, WI), dtype=torch.uint8, device=a.device)
        native.check(lib.node_b200_convs2_prepare(native.ptr(buf), C, HI, WI, native.ptr(conv.weight), native.ptr(down.weight),
                                                  native.ptr(norm.weight), native.ptr(norm.bias), native.stream_ptr()), 'convs2_prepare')
        ent = _convs2_ws[key] = (buf, ver, [weakref.ref(p) for p in live])
    if _needs_grad(a, conv.weight, down.weight):
        return caller_grad.ResHead.apply(a, conv.weight, down.weight, ent[0])
    c = torch.empty((N, C, HO, WO), dtype=a.dtype, device=a.device)
    sc = torch.empty_like(c)
    native.check(lib.node_b200_convs2_forward(native.ptr(ent[0]), native.ptr(a), native.ptr(c), native.ptr(sc), N, C, HI, WI,
                                              native.stream_ptr()), 'convs2_forward')
    global launches
    launches += 1
    return c, sc


# ---- fused stem: relu(norm(conv(x))) for the first Conv2d(CIN, 64, 3, 1) of a downsampler (csrc/caller_ops.cu) -------------

def _stem_ok(conv, norm, x):
    if not enabled():
        return False
    if not (isinstance(conv, nn.Conv2d) and isinstance(norm, nn.GroupNorm) and norm.weight is not None and conv.bias is not None):
        return False
    if (conv.out_channels, conv.kernel_size, conv.stride, conv.padding, conv.dilation, conv.groups) != (64, (3, 3), (1, 1), (0, 0), (1, 1), 1):
        return False
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and norm.num_groups == 32 and norm.num_channels == 64):
        return False
    if torch.is_grad_enabled() and x.requires_grad:          # dL/dx wanted (PGD attack): the modules' own ops
        return False
    if _needs_grad(conv.weight, conv.bias, norm.weight, norm.bias) and not caller_grad.enabled():
        return False
    return (int(x.shape[1]), int(x.shape[2]), int(x.shape[3])) in ((3, 32, 32), (1, 28, 28)) and conv.in_channels == x.shape[1]


@native.on_device_of(2)
def stem_gn_relu(conv, norm, x):
    """relu(norm(conv(x))): one CUDA pass when served and no gradient is needed, the modules' own ops otherwise."""
    if not _stem_ok(conv, norm, x):
        return group_norm_relu(norm, conv(x))
    if _needs_grad(conv.weight, conv.bias, norm.weight, norm.bias):
        return caller_grad.Stem.apply(x, conv.weight, conv.bias, norm.weight, norm.bias, float(norm.eps))
    x = x.contiguous()
    N, CIN, HIN, WIN = (int(v) for v in x.shape)
    out = torch.empty((N, 64, HIN - 2, WIN - 2), dtype=x.dtype, device=x.device)
    native.check(native.lib().node_b200_stem_gn_relu(native.ptr(x), native.ptr(conv.weight), native.ptr(conv.bias), native.ptr(norm.weight),
                                                     native.ptr(norm.bias), native.ptr(out), N, CIN, HIN, WIN, float(norm.eps),
                                                     native.stream_ptr()), 'stem_gn_relu')
    global launches
    launches += 1
    return out


# ---- fused classifier head: GroupNorm -> ReLU -> global average pool (-> Linear) (csrc/caller_ops.cu k_head) ---------------

@native.on_device_of(1)
def head(seq, x, out=None):
    """FCClassifier.module (model.py:231-250) applied to x: one CUDA pass when the layer list is exactly
    [GroupNorm(32, 64), ReLU, AdaptiveAvgPool2d(1), (Dropout in eval mode,) Flatten, Linear(64, k) or an empty Sequential]
    and no gradient is needed; run_sequential otherwise. `out` (contiguous [N, n_out] fp32 on x's device, e.g. a slice of
    the features[tol, T, N, 64] buffer of node_b200.retrieval.FeatureStore) receives the result in place."""
    mods = [m for m in seq.children() if not (isinstance(m, nn.Dropout) and not m.training)]
    ok = (len(mods) == 5 and isinstance(mods[0], nn.GroupNorm) and isinstance(mods[1], nn.ReLU)
          and isinstance(mods[2], nn.AdaptiveAvgPool2d) and mods[2].output_size in (1, (1, 1)) and type(mods[3]).__name__ == 'Flatten'
          and (isinstance(mods[4], nn.Linear) or (type(mods[4]) is nn.Sequential and len(mods[4]) == 0)))
    ok = ok and enabled()
    norm = mods[0] if ok else None
    if ok:
        lin = mods[4] if isinstance(mods[4], nn.Linear) else None
        ok = (norm.weight is not None and norm.num_groups == 32 and norm.num_channels == 64 and x.is_cuda and x.dtype == torch.float32
              and x.dim() == 4 and x.shape[1] == 64 and x.shape[0] > 0
              and (lin is None or (lin.in_features == 64 and lin.out_features <= 256 and lin.weight.dtype == torch.float32))
              and not (torch.is_grad_enabled() and (x.requires_grad or norm.weight.requires_grad
                                                    or (lin is not None and lin.weight.requires_grad))))
    if not ok:
        res = run_sequential(seq, x)
        if out is not None:
            out.copy_(res)
            return out
        return res
    x = x.contiguous()
    N, HW = int(x.shape[0]), int(x.shape[2] * x.shape[3])
    n_out = lin.out_features if lin is not None else 64
    if out is None:
        out = torch.empty((N, n_out), dtype=x.dtype, device=x.device)
    elif not (out.is_contiguous() and tuple(out.shape) == (N, n_out) and out.dtype == x.dtype and out.device == x.device):
        raise ValueError('head: `out` must be a contiguous [%d, %d] %s tensor on %s' % (N, n_out, x.dtype, x.device))
    lw = native.ptr(lin.weight.contiguous()) if lin is not None else None
    lb = native.ptr(lin.bias) if lin is not None and lin.bias is not None else None
    native.check(native.lib().node_b200_head(native.ptr(x), native.ptr(norm.weight), native.ptr(norm.bias), lw, lb, native.ptr(out),
                                             N, 64, HW, n_out, float(norm.eps), native.stream_ptr()), 'head')
    global launches
    launches += 1
    return out
