import os, sys, time
sys.path.insert(0, 'neural-ode-features_b200'); sys.path.insert(0, '.')
import torch, torch.nn.functional as F
import __graft_entry__ as e; e.build()
from node_b200 import models, solver
torch.backends.cudnn.allow_tf32 = False
dev = 'cuda'
for B in (1, 2, 4):
    for mode in ('1', '0'):
        os.environ['NODE_B200_ADJOINT_STEP'] = mode
        torch.manual_seed(0)
        net = models.ODENet(3, n_filters=64, downsample='residual', tol=1e-3, adjoint=True).to(dev).train()
        x = torch.rand(B, 3, 32, 32, device=dev); y = torch.randint(0, 10, (B,), device=dev)
        loss = F.cross_entropy(net(x), y); nf = net.nfe(reset=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        loss.backward()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        st = solver.last_stats
        g = net.odeblock.odefunc.conv1._layer.weight.grad
        print('B', B, 'mode', mode, 'nfe_f', nf, 'nfe_b', net.nfe(reset=True), 'accept', st.get('n_accept'), 'reject', st.get('n_reject'), 'route', st.get('route'),
              'bwd ms %.2f' % (dt * 1e3), 'gnorm %.6e' % float(g.norm()), 'dt trace', [round(float(v), 4) for v in st['trace']['dt'][:8]])
