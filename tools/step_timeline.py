"""Tuning aid: conv-job timeline of CTA 0 of the step kernel (build with NVCC_EXTRA=-DNODE_STEP_DEBUG)."""
import ctypes, sys, os
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'neural-ode-features_b200'), ROOT]
import __graft_entry__ as entry
entry.build()
from node_b200 import models, native, solver
torch.manual_seed(0)
net = models.ODENet(3, n_filters=64, downsample='residual', tol=1e-3).eval().cuda()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = torch.rand(B, 3, 32, 32, device='cuda')
lib = native.lib()
lib.node_b200_step_debug_read2.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
buf2 = (ctypes.c_longlong * 1024)()
with torch.no_grad():
    h0 = net.downsample(x)
    for _ in range(2):
        net.odeblock(h0)
    torch.cuda.synchronize()
    lib.node_b200_step_debug_read2(buf2, 1024, 1)
    ph = (ctypes.c_longlong * 32)()
    lib.node_b200_step_phase_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.node_b200_step_phase_read(ph, 1)
    solver._step_guess.clear()
    net.odeblock(h0)
    torch.cuda.synchronize()
lib.node_b200_step_debug_read2(buf2, 1024, 0)
w2 = np.array(buf2[:], dtype=np.int64).reshape(2, 256, 2)
print('leader waits per job (accumulated over the launches of one solve, /#launches for one): wfull', np.median(w2[0, :60, 0]), 'wfree', np.median(w2[0, :60, 1]))
buf = (ctypes.c_longlong * 2048)()
lib.node_b200_step_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
print('rc', lib.node_b200_step_debug_read(buf, 2048))
a = np.array(buf[:], dtype=np.int64).reshape(2, 256, 4)
t0 = a[a > 0].min()
for s in range(2):
    rows = a[s]
    n = int((rows[:, 0] > 0).sum())
    d = rows[:n] - t0
    print('slot', s, 'jobs', n)
    turn = d[:, 1] - d[:, 0]; issue = d[:, 2] - d[:, 1]; acc = d[:, 3] - d[:, 2]; work = np.diff(d[:, 0], prepend=0) - 0
    print(' median clk: wait-turn %d  issue %d  wait-acc-after-issue %d  | publish->acc %d | acc->next publish (worker phase) %d' % (
        np.median(turn), np.median(issue), np.median(acc), np.median(d[:, 3] - d[:, 0]), np.median(d[1:, 0] - d[:-1, 3])))
    for j in range(min(n, 14)):
        print('  job %3d  pub %8d turn %8d issued %8d acc %8d' % (j, d[j, 0], d[j, 1], d[j, 2], d[j, 3]))

lib.node_b200_step_phase_read(ph, 0)
names = ['make_tb', 'stage_in', 'gn1', 'apply1', 'wait conv1', 'read c1', 'gn2', 'apply2', 'wait conv2', 'read c2', 'gn3', 'affine+k store',
         'error norm']
st = solver.last_stats
nst = st['n_accept'] + st['n_reject']
for s_ in range(2):
    v = np.array(ph[s_ * 16:s_ * 16 + 13], dtype=np.float64)
    print('   leader clocks inside the MMA issue blocks per step launch: %.0f' % (ph[s_ * 16 + 13] / nst))
    print('slot', s_, 'thread 32 of CTA 0, clocks per step launch (%d launches), total %.0f' % (nst, v.sum() / nst))
    for n_, x_ in zip(names, v):
        print('   %-16s %9.0f  %5.1f%%' % (n_, x_ / nst, 100 * x_ / v.sum()))
