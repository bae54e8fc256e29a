"""CPU, world_size 2 over gloo: the host-side sharding logic of SURVEY 8(e). Each rank owns half
of the batch and contributes float64 partial sums to every norm; the step sequence must be the
one the single-shard solve takes on the whole batch."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import dopri5_port, odefunc_port
    g = dict(np.load(os.path.join(GOLDEN, 'cifar_res_n8.npz')))
    p = {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('p.')}
    h0, t, tol = torch.from_numpy(g['h0']), torch.from_numpy(g['t']), float(g['tol'])
    shard = h0.chunk(world)[rank]

    def reduce(ssq, n):
        v = torch.tensor([float(ssq), float(n)], dtype=torch.float64)
        dist.all_reduce(v)
        return v[0], v[1]

    tr = dopri5_port.Trace()
    torch.set_num_threads(2)
    with torch.no_grad():
        out = dopri5_port.dopri5_solve(lambda a, b: odefunc_port.odefunc_forward(p, a, b), shard, t, tol, tol,
                                       trace=tr, norm_reduce=reduce)
    full = torch.from_numpy(g['out']).chunk(world, dim=1)[rank]
    ret[rank] = dict(acc=[s[2] for s in tr.steps], dt=[s[1] for s in tr.steps], nfe=tr.nfe,
                     err=float((out - full).abs().max() / full.abs().max()))
    dist.destroy_process_group()


def test_two_rank_sharded_solve_takes_the_global_step_sequence():
    g = dict(np.load(os.path.join(GOLDEN, 'cifar_res_n8.npz')))
    ret = mp.Manager().dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    for rank in (0, 1):
        r = ret[rank]
        assert r['acc'] == list(g['tr_acc'])
        np.testing.assert_allclose(r['dt'], g['tr_dt'], rtol=1e-5)
        assert r['nfe'] == int(g['nfe'])
        assert r['err'] < 2e-5
    assert ret[0]['dt'] == ret[1]['dt']          # identical controller decisions on every rank
