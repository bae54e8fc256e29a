"""Tuning aid: one cfg3 training step (forward + CE + odeint_adjoint backward + SGD) between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file ... python tools/train_profile.py [batch]
and CUDA-event timing of its forward / backward halves without the profiler."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'neural-ode-features_b200'), ROOT]
import torch
import __graft_entry__ as entry
entry.build()
import bench
from node_b200 import solver

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device('cuda', 0)
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
net = bench.build_model(dev, adjoint=True).train()
opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
x = torch.rand(B, 3, 32, 32, device=dev)
y = torch.randint(0, 10, (B,), device=dev)


def step(ev=None):
    opt.zero_grad(set_to_none=True)
    if ev: ev[0].record()
    loss = torch.nn.functional.cross_entropy(net(x), y)
    if ev: ev[1].record()
    loss.backward()
    if ev: ev[2].record()
    opt.step()
    if ev: ev[3].record()


for _ in range(3):
    step()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
step(ev)
torch.cuda.synchronize()
print('batch %d: forward %.2f ms  backward %.2f ms  optimizer %.2f ms' % (B, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])))
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
