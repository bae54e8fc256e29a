"""Callers of the hot path: GroupNorm -> ReLU of the downsamplers and the classifier head
(reference model.py:119-178, 231-250, 268-271) as one memory-bound CUDA pass (csrc/caller_ops.cu).

Used by `node_b200.models` for inference (no autograd graph); with gradients enabled the modules run their own
PyTorch ops, exactly as in the reference. The module tree - and so the state_dict keys - is unchanged."""
import torch
import torch.nn as nn

from . import native


def _fusable(norm, x):
    if not isinstance(norm, nn.GroupNorm) or norm.weight is None or norm.bias is None:
        return False
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() >= 3 and norm.weight.is_cuda):
        return False
    if torch.is_grad_enabled() and (x.requires_grad or norm.weight.requires_grad or norm.bias.requires_grad):
        return False
    cell = (x.shape[1] // norm.num_groups) * (x.numel() // (x.shape[0] * x.shape[1]))
    return x.shape[1] == norm.num_channels and 0 < cell <= 4096 and x.shape[0] > 0


def group_norm_relu(norm, x, relu=True):
    """relu(norm(x)) (or norm(x)) for an nn.GroupNorm `norm`; one CUDA pass when no gradient is needed."""
    if not _fusable(norm, x):
        y = norm(x)
        return torch.relu(y) if relu else y
    x = x.contiguous()
    y = torch.empty_like(x)
    N, C = int(x.shape[0]), int(x.shape[1])
    err = native.lib().node_b200_groupnorm_relu(native.ptr(x), native.ptr(y), native.ptr(norm.weight), native.ptr(norm.bias),
                                               N, C, int(norm.num_groups), x.numel() // (N * C), float(norm.eps),
                                               1 if relu else 0, native.stream_ptr())
    native.check(err, 'groupnorm_relu')
    return y


def run_sequential(seq, x):
    """nn.Sequential.forward with every (GroupNorm, ReLU) pair fused."""
    mods = list(seq.children())
    i = 0
    while i < len(mods):
        m = mods[i]
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
