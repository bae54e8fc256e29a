"""Tuning aid: per-phase clocks of worker thread 0 and of the issuer of CTA 0 of the dense step kernel.
Needs a debug build: tools/build_variant.sh dbg -DNODE_STEP8_DEBUG ; NODE_B200_LIB=<that .so> python tools/step8_timeline.py [batch]"""
import ctypes, sys, os
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'neural-ode-features_b200'), ROOT]
from node_b200 import models, native, solver
torch.manual_seed(0)
net = models.ODENet(3, n_filters=64, downsample='residual', tol=1e-3).eval().cuda()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
x = torch.rand(B, 3, 32, 32, device='cuda')
lib = native.lib()
lib.node_b200_step8_phase_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
ph = (ctypes.c_longlong * 32)()
with torch.no_grad():
    h0 = net.downsample(x)
    for _ in range(2):
        net.odeblock(h0)
    torch.cuda.synchronize()
    lib.node_b200_step8_phase_read(ph, 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); net.odeblock(h0); b.record()
    torch.cuda.synchronize()
lib.node_b200_step8_phase_read(ph, 0)
st = solver.last_stats
nst = st['n_accept'] + st['n_reject']
v = np.array(ph[:], dtype=np.float64) / nst
names = {0: 'W1 stage-in loads+combine', 1: 'W1 gn1 (quad)', 2: 'W1 A write + publish', 15: 'other (loop, before wait)', 3: 'W2 wait conv1',
         4: 'W2 read c1', 5: 'W2 gn2', 6: 'W2 A write + publish', 7: 'W3 wait conv2', 8: 'W3 read c2', 9: 'W3 gn3', 10: 'W3 affine + k store',
         11: 'W3 error norm'}
tot = sum(v[i] for i in names)
print('ODE block %.3f ms, %d step launches; worker thread 0 of CTA 0: %.0f clocks per step launch' % (a.elapsed_time(b), nst, tot))
for i in (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 15):
    print('   %-28s %9.0f  %5.1f%%' % (names[i], v[i], 100 * v[i] / tot))
print('issuer: between jobs %.0f  wait A ready %.0f  (tap loop) issue+other %.0f  wait own weights %.0f  wait peer weights %.0f' % (v[16], v[17], v[18], v[19], v[20]))
print('producer (CTA 0): request+loop %.0f  wait ring slot free %.0f' % (v[24], v[25]))
