"""Batch sharding across the GPUs of one box.

ODEfunc has no cross-sample coupling (GroupNorm statistics are per sample), so the batch shards
with no data-path collective; the only cross-sample quantities are the error norm and the
initial-step norms (misc.py:123-136,155-156), which are means over the WHOLE batch. Each rank
therefore reduces its shard to float64 partial sums on the device, the sums (<= 16 doubles) are
all-reduced over NCCL/NVLink, and every rank runs the identical device-side controller on
identical inputs - the accepted-step sequence is the same on every rank and equal to the
single-GPU run at the same global batch.

enable(group) switches the solver into sharded mode for this process; disable() switches back.

Training (SURVEY 8e "training collective"): with the loss a mean over the LOCAL shard, the global gradient is the average
of the ranks' gradients. The adjoint already integrates adj_params as a GLOBAL sum (its error norm needs the global value,
adjoint.py:35-55 never reads it back), so after backward() the ODE-block parameters hold the SUM over ranks while every
other parameter holds its local gradient. sync_gradients(model) brings both to the global average with ONE bucketed
all-reduce of the non-ODE gradients (158,730 floats for the CIFAR `residual` net); do not wrap the model in
DistributedDataParallel as well - it would average the already reduced ODE-block gradients a second time.
"""
import torch

_group = None


_peer = False


def enable(process_group=None, peer_reduce=None):
    """Switch the solver into sharded mode. With `peer_reduce` (default: on for CUDA ranks of one box, NODE_B200_PEER_REDUCE=0
    switches it off) the fused route's sums are all-reduced over NVLink peer memory INSIDE the kernel that folds them
    (csrc/peer_reduce.cu) instead of a host-launched ncclAllReduce per attempted step: the ranks exchange the CUDA IPC handles
    of their exchange buffers once, here."""
    import os
    import torch.distributed as dist
    global _group, _peer
    if not dist.is_initialized():
        raise RuntimeError('torch.distributed is not initialised')
    _group = process_group if process_group is not None else dist.group.WORLD
    _peer = False
    if peer_reduce is None:
        peer_reduce = os.environ.get('NODE_B200_PEER_REDUCE', '1') != '0'
    world = dist.get_world_size(_group)
    if peer_reduce and torch.cuda.is_available() and dist.get_backend(_group) == 'nccl' and 1 < world <= 8:
        _peer = _setup_peers(world, dist.get_rank(_group))


def _setup_peers(world, rank):
    """Exchange the IPC handles of the ranks' exchange buffers; every rank reports whether it could map all peers and the
    mode is used only if ALL could (the decision must be identical everywhere: it changes the collective sequence)."""
    import ctypes
    import torch.distributed as dist
    from . import native
    lib = native.lib()
    dev = torch.device('cuda', torch.cuda.current_device())
    handle = (ctypes.c_ubyte * 64)()
    ok = lib.node_b200_peer_alloc(ctypes.cast(handle, ctypes.c_void_p)) == 0
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=_group)
    if ok:
        blob = torch.cat(gathered).cpu().numpy().tobytes()
        buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
        ok = lib.node_b200_peer_open(world, rank, ctypes.cast(buf, ctypes.c_void_p)) == 0
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=_group)
    if int(flag.item()) != 1:
        lib.node_b200_peer_close()
        return False
    dist.barrier(group=_group)              # every rank has zeroed and mapped the buffers before anybody writes
    return True


def peer_active():
    """True when the fused route reduces its sums over peer memory (no NCCL call per attempted step)."""
    return _group is not None and _peer


def disable():
    global _group, _peer
    if _peer:
        from . import native
        native.lib().node_b200_peer_close()
    _group = None
    _peer = False


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


def world_size():
    import torch.distributed as dist
    return dist.get_world_size(_group)


def ode_parameters(model):
    """Parameters whose gradient comes out of odeint_adjoint already summed over the ranks: those of every module the
    fused / generic solver integrates (`ODEBlock.odefunc`, model.py:352-360)."""
    seen = {}
    for m in model.modules():
        f = getattr(m, 'odefunc', None)
        if f is not None:
            for q in f.parameters():
                seen[id(q)] = q
    return seen


def sync_gradients(model):
    """Global-average gradients on every rank (call between backward() and optimizer.step()); returns the number of
    floats that crossed the link. No-op when sharding is off."""
    if _group is None:
        return 0
    world = world_size()
    ode = ode_parameters(model)
    rest = [q for q in model.parameters() if q.grad is not None and id(q) not in ode]
    sent = 0
    if rest:
        flat = torch.cat([q.grad.reshape(-1) for q in rest])          # one bucket: < 1 MB for every net of the reference
        all_reduce_sum(flat)
        flat.div_(world)
        sent = flat.numel()
        o = 0
        for q in rest:
            n = q.grad.numel()
            q.grad.copy_(flat[o:o + n].view_as(q.grad))
            o += n
    for q in ode.values():
        if q.grad is not None:
            q.grad.div_(world)                                        # already the sum over ranks (adjoint collective)
    return sent
