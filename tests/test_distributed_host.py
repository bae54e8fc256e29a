"""CPU, world_size 2 over gloo: the host-side logic of node_b200.distributed itself (the solver needs a GPU; what it asks
of this module does not): the global element count that normalises the batch-global error norm - including the uneven /
changing batch that a cached count would get wrong (the collective sequences of the ranks must never diverge) - and the
training collective sync_gradients: non-ODE gradients averaged in one bucket, ODE-block gradients, which the
adjoint has already summed over the ranks, only divided."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


class _Func(nn.Module):
    def __init__(self):
        super().__init__()
        self.lin = nn.Linear(3, 3)


class _Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.odefunc = _Func()


class _Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.down = nn.Linear(4, 3)
        self.odeblock = _Block()
        self.head = nn.Linear(3, 2)
        self.unused = nn.Linear(2, 2)          # never gets a gradient: must be skipped, not crash


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from node_b200 import distributed as nd
    out = {}
    assert nd.group() is None and nd.sync_gradients(_Net()) == 0           # off: no collective
    nd.enable()
    # shard sizes 128/128, then 128/72 (an uneven last batch): rank 0's local size does not change, the global count does
    out['numel'] = [nd.global_numel(128 * 7, 'cpu'), nd.global_numel((128 if rank == 0 else 72) * 7, 'cpu'),
                    nd.global_numel(128 * 7, 'cpu')]
    v = torch.tensor([1.0 + rank, 10.0 * (rank + 1)], dtype=torch.float64)
    nd.all_reduce_sum(v)
    out['sum'] = v.tolist()
    torch.manual_seed(0)
    net = _Net()
    for i, q in enumerate(net.parameters()):
        q.grad = None
    ode = nd.ode_parameters(net)
    out['n_ode'] = len(ode)
    for i, (name, q) in enumerate(net.named_parameters()):
        if name.startswith('unused'):
            continue
        if id(q) in ode:
            q.grad = torch.full_like(q, 6.0 + i)            # what the adjoint leaves: the SUM over ranks, identical on every rank
        else:
            q.grad = torch.full_like(q, float(rank + 1) * (i + 1))   # local gradients
    sent = nd.sync_gradients(net)
    out['sent'] = sent
    out['grads'] = {name: (None if q.grad is None else float(q.grad.flatten()[0])) for name, q in net.named_parameters()}
    nd.disable()
    ret[rank] = out
    dist.destroy_process_group()


def test_distributed_host_logic_two_ranks():
    ret = mp.Manager().dict()
    port = 31500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    for rank in (0, 1):
        r = ret[rank]
        assert r['numel'] == [2 * 128 * 7, (128 + 72) * 7, 2 * 128 * 7]      # never served from a stale cache
        assert r['sum'] == [3.0, 30.0]
        assert r['n_ode'] == 2
        assert r['sent'] == 4 * 3 + 3 + 3 * 2 + 2                             # down + head, one bucket
        g = r['grads']
        names = list(g)
        for i, name in enumerate(names):
            if name.startswith('unused'):
                assert g[name] is None
            elif name.startswith('odeblock'):
                assert g[name] == (6.0 + i) / 2                                # sum over ranks -> average
            else:
                assert g[name] == 1.5 * (i + 1)                                # (1 + 2) / 2 times (i + 1)
    assert ret[0]['grads'] == ret[1]['grads']
