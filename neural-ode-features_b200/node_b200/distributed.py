"""Batch sharding across the GPUs of one box.

ODEfunc has no cross-sample coupling (GroupNorm statistics are per sample), so the batch shards
with no data-path collective; the only cross-sample quantities are the error norm and the
initial-step norms (misc.py:123-136,155-156), which are means over the WHOLE batch. Each rank
therefore reduces its shard to float64 partial sums on the device, the sums (<= 16 doubles) are
all-reduced over NCCL/NVLink, and every rank runs the identical device-side controller on
identical inputs - the accepted-step sequence is the same on every rank and equal to the
single-GPU run at the same global batch.

enable(group) switches the solver into sharded mode for this process; disable() switches back.
"""
import torch

_group = None


def enable(process_group=None):
    import torch.distributed as dist
    global _group
    if not dist.is_initialized():
        raise RuntimeError('torch.distributed is not initialised')
    _group = process_group if process_group is not None else dist.group.WORLD


def disable():
    global _group
    _group = None


def group():
    return _group


def rank():
    import torch.distributed as dist
    return dist.get_rank(_group)


def all_reduce_sum(t):
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_group)


def global_numel(local_numel, device):
    """Sum of the shard sizes over the group: one tiny all-reduce per call, by EVERY rank. Never cached - a cache keyed by
    the local size would skip the collective on the ranks whose shard size did not change when the global batch did
    (uneven last batch), and the ranks' collective sequences would diverge."""
    import torch.distributed as dist
    v = torch.tensor([int(local_numel)], dtype=torch.int64, device=device)
    dist.all_reduce(v, op=dist.ReduceOp.SUM, group=_group)
    return int(v.item())
