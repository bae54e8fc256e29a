"""Bring-up aid for the dense 8x8 engine: one evaluation / one solve on the golden cifar_res_n8 case against the golden
reference values. Run under compute-sanitizer when it faults."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'neural-ode-features_b200'), ROOT, os.path.join(ROOT, 'tests')]
import __graft_entry__ as entry
entry.build()
from node_b200 import solver, odeint
from conftest import load_odefunc
g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'cifar_res_n8.npz')))
dev = torch.device('cuda')
func = load_odefunc(g, dev)
h0 = torch.from_numpy(g['h0']).to(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = (n + 7) // 8
hh = h0.repeat(reps, 1, 1, 1)[:n].contiguous()
ref = torch.from_numpy(g['f037']).repeat(reps, 1, 1, 1)[:n]
k = solver.odefunc_forward(func, 0.37, hh)
torch.cuda.synchronize()
err = float((k.cpu() - ref).abs().max() / ref.abs().max())
print('eval N=%d rel err %.3e' % (n, err))
with torch.no_grad():
    out = odeint(func, hh, torch.from_numpy(g['t']).to(dev), rtol=1e-3, atol=1e-3, method='dopri5')
torch.cuda.synchronize()
refo = torch.from_numpy(g['out']).repeat(1, reps, 1, 1, 1)[:, :n]
print('solve rel err %.3e nfe %d (golden %d) accepted %s' % (float((out.cpu() - refo).abs().max() / refo.abs().max()), solver.last_stats['nfe'], int(g['nfe']), list(solver.last_stats['trace']['accepted'])))
