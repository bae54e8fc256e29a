"""N-GPU parity of the sharded solve: every rank integrates its shard of one global batch with the
batch-global error norm all-reduced over NCCL; rank 0 also integrates the WHOLE batch alone. The step sequence must
be identical and the outputs equal to fp32 reduction-order noise. Run under torchrun; prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'neural-ode-features_b200'), ROOT]
import torch
import torch.distributed as dist

import __graft_entry__ as entry
entry.build()
from node_b200 import models, solver, distributed as nd

rank, local, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
torch.backends.cudnn.allow_tf32 = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 96 * world
torch.manual_seed(0)
net = models.ODENet(3, n_filters=64, downsample='residual', tol=1e-3).eval().to(dev)
x = torch.rand(B, 3, 32, 32, generator=torch.Generator().manual_seed(7)).to(dev)
t = torch.tensor([0., 0.3, 1.0], device=dev)
res = {}
with torch.no_grad():
    h0 = net.downsample(x)
    per = B // world
    nd.enable()
    out = solver.odeint(net.odeblock.odefunc, h0[rank * per:(rank + 1) * per].contiguous(), t, rtol=1e-3, atol=1e-3, method='dopri5')
    st = dict(solver.last_stats)
    nd.disable()
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    traces = [None] * world
    dist.all_gather_object(traces, dict(nfe=st['nfe'], acc=list(map(int, st['trace']['accepted'])), dt=list(map(float, st['trace']['dt']))))
    if rank == 0:
        full = solver.odeint(net.odeblock.odefunc, h0, t, rtol=1e-3, atol=1e-3, method='dopri5')
        s1 = dict(solver.last_stats)
        sh = torch.cat(gathered, dim=1)
        rel = float((sh - full).abs().max() / full.abs().max())
        same = all(tr['nfe'] == s1['nfe'] and tr['acc'] == list(map(int, s1['trace']['accepted'])) for tr in traces)
        dtdev = max(abs(a - b) / b for tr in traces for a, b in zip(tr['dt'], map(float, s1['trace']['dt'])))
        res = dict(world=world, global_batch=B, nfe=s1['nfe'], accepted=list(map(int, s1['trace']['accepted'])),
                   identical_step_sequence_on_all_ranks=bool(same), max_rel_dt_deviation=dtdev, max_rel_output_deviation=rel,
                   route=st['route'])
        print(json.dumps(res))
        assert same and dtdev < 1e-6 and rel < 1e-5, res
dist.barrier()
dist.destroy_process_group()
