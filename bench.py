"""bench.py - headline benchmark of the B200-native ODE-Net dopri5 hot path.

Metric (BASELINE.json): CIFAR-10 ODENet dopri5 (tol 1e-3) forward, images/s, synthetic 32x32
batches, random-init weights, fp32 contract (3xTF32 tensor-core convolution).  A "step" is one
ODENet forward over one per-GPU batch; weak scaling (per-GPU batch fixed, the error norm is
all-reduced over the GLOBAL batch every attempted step).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port), host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, 'neural-ode-features_b200'), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

FLOP_PER_IMG_PER_NFE_PER_PIXEL = 147456          # 2 convs x 2*9*64*64, algorithmic
TOL = 1e-3


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p['hbm_gbs'], bf16=p['bf16_tflops'], bf16_sus=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                    src='measured')
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src='fallback')


class ClockSampler(object):
    Q = 'timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                       '-lms', '20'], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def mark(self):
        """Wall-clock window of the timed region: only samples inside it count (the sampler itself starts earlier, before
        the warm-up, because nvidia-smi needs ~100 ms to produce its first line)."""
        self.t0 = time.time()

    def stop(self):
        if self.p is None:
            return None
        t1 = time.time()
        time.sleep(0.03)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(', ') for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        import datetime
        t0 = getattr(self, 't0', 0.0)
        sm, mx, reasons, n_all = [], 0.0, set(), 0
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                ts = datetime.datetime.strptime(r[0].strip(), '%Y/%m/%d %H:%M:%S.%f').timestamp()
                clk, cmax = float(r[1]), float(r[2])
            except (ValueError, IndexError):
                continue
            n_all += 1
            mx = max(mx, cmax)
            if not (t0 - 0.02 <= ts <= t1 + 0.02):
                continue
            sm.append(clk)
            for n, v in zip(names, r[4:8]):
                if v.strip().lower().startswith('active'):
                    reasons.add(n)
        if not sm:
            return None
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm), samples_total=n_all)


def build_model(device, adjoint=False, dropout=0):
    from node_b200 import models
    torch.manual_seed(0)
    net = models.ODENet(3, n_filters=64, downsample='residual', tol=TOL, adjoint=adjoint, dropout=dropout).eval()
    return net.to(device)


class _PortAdjoint(torch.autograd.Function):
    """The oracle's restatement of odeint_adjoint (adjoint.py:7-133) behind autograd, for the reference arms only."""

    @staticmethod
    def forward(ctx, func, t, tol, y0, *params):
        from oracle import dopri5_port
        ctx.func, ctx.tol, ctx.params = func, tol, params
        with torch.no_grad():
            ys = dopri5_port.dopri5_solve(lambda a, b: func(a, b), y0, t, tol, tol)
        ctx.save_for_backward(t, ys)
        return ys

    @staticmethod
    def backward(ctx, g):
        from oracle import dopri5_port
        t, ys = ctx.saved_tensors
        gy, gt, gp = dopri5_port.adjoint_backward(lambda a, b: ctx.func(a, b), ctx.params, t, ys, g, ctx.tol, ctx.tol)
        outs, o = [], 0
        for q in ctx.params:
            outs.append(gp[o:o + q.numel()].view_as(q))
            o += q.numel()
        return (None, None, None, gy) + tuple(outs)


def reference_model(device, in_ch=3, downsample='residual', adjoint=False, train=False):
    """The reference's own op sequence on `device`: node_b200.models mirrors model.py (same modules, same state dict), its
    odeint replaced by the oracle's torch restatement of the pinned torchdiffeq (oracle/dopri5_port.py) - ATen kernels only,
    none of this repo's CUDA code. /root/reference does not travel to the GPU box; tools/make_golden.py pins the
    restatement bit for bit against it."""
    from node_b200 import models
    from oracle import dopri5_port
    torch.manual_seed(0)
    net = models.ODENet(in_ch, n_filters=64, downsample=downsample, tol=TOL, adjoint=adjoint)
    net = net.train() if train else net.eval()
    net = net.to(device)
    func = net.odeblock.odefunc
    if adjoint:
        net.odeblock.odeint = lambda f, y0, t, **kw: _PortAdjoint.apply(f, t, kw['rtol'], y0, *tuple(f.parameters()))
    else:
        net.odeblock.odeint = lambda f, y0, t, **kw: dopri5_port.dopri5_solve(lambda a, b: f(a, b), y0, t, kw['rtol'], kw['atol'])
    return net


def timed_host(fn, warmup, reps):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(reps):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts)


def cpu_forward_rate(sample, min_seconds, threads):
    """The reference's CPU path (torch-CPU restatement = same ATen kernels, all host threads)."""
    torch.set_num_threads(threads)
    net = reference_model('cpu')
    x = torch.rand(sample, 3, 32, 32)
    times = []
    with torch.no_grad():
        net(x)                                   # warm-up
        t_all = time.perf_counter()
        while True:
            t0 = time.perf_counter()
            net(x)
            times.append(time.perf_counter() - t0)
            if (time.perf_counter() - t_all >= min_seconds and len(times) >= 2) or len(times) >= 20:
                break
    return sample / statistics.median(times), times


def run_reference(args):
    """--impl reference: rank 0 alone times the reference's CPU path on the benchmark's own configuration (the whole per-GPU
    batch per step); other ranks exit."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sample = args.batch if args.cpu_sample <= 0 else min(args.batch, args.cpu_sample)
    net = reference_model('cpu')
    x = torch.rand(sample, 3, 32, 32)
    with torch.no_grad():
        for _ in range(args.warmup):
            net(x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            net(x)
        dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    what = 'the whole per-GPU batch' if sample == args.batch else 'a %d-image sample' % sample
    line = dict(impl='reference', metric='CIFAR-10 ODENet dopri5 forward throughput', value=v, unit='images/s', n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f32', data='synthetic',
                config=workload_config(args, sample_note='reference arm: %s (%d images) per step on the host CPU' % (what, sample)),
                cpu_baseline=dict(value=v, unit='images/s', cores=threads, kind='port',
                                  sample='%d images per step, %d steps, torch-CPU restatement of the reference (oracle/)' % (sample, args.steps)),
                e2e=dict(value=v, unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def workload_config(args, sample_note=None):
    c = dict(workload='cfg2: CIFAR-10 ODENet(3, n_filters=64, downsample=residual, tol=1e-3).eval() forward, '
                      'synthetic 32x32 batches', per_gpu_batch=args.batch, global_batch=args.batch * args.gpus,
             solver='dopri5 rtol=atol=1e-3', conv_mode=os.environ.get('NODE_B200_CONV', 'f16x3'),
             downsample_classifier='own CUDA kernels for the stem (conv+GN+ReLU), the ResBlock heads (3x3 s2 + 1x1 s2, tcgen05) and '
                                   'tails (GN->ReLU->conv3x3->add, tcgen05) and the remaining GroupNorm->ReLU pairs; pooling / linear: PyTorch',
             batch_note='per-GPU batch sized to whole rounds of the persistent step kernel: 74 CTA pairs x 2 CTAs x 2 virtual slots x 4 images = 1184 images per round, %d rounds' % max(1, args.batch // 1184) if args.batch % 1184 == 0 else 'per-GPU batch %d (not a multiple of the 1184-image round of the persistent step kernel)' % args.batch,
             parallelism='dp%d batch shard, error-norm allreduce' % args.gpus,
             l2='inputs larger than L2 (state %d MB per tensor, ~10 live tensors)' % (args.batch * 64 * 64 * 4 // 2 ** 20))
    if sample_note:          # the reference arm: the oracle port on the host CPU, none of the kernels above
        c['sample'] = sample_note
        c['conv_mode'] = 'ATen CPU fp32'
        c['downsample_classifier'] = 'ATen CPU fp32'
        c.pop('batch_note', None)
    return c


def rk_roofline(device, pk):
    """HBM roofline of the standalone RK kernels (generic route): stage-6 combination and error norm."""
    import ctypes
    from node_b200 import native
    import numpy as np
    lib = native.lib()
    n = 32 * 1024 * 1024                                  # 128 MiB per tensor, 9 tensors >> L2
    bufs = [torch.randn(n, device=device) for _ in range(9)]
    L = native.layout()
    ctl = torch.zeros(L['sizeof'], dtype=torch.uint8, device=device)
    native.check(lib.node_b200_ctl_init(native.ptr(ctl), native.F32, 1, native.host_f64([TOL]),
                                        native.host_f64([TOL]), native.host_i64([n]),
                                        0.9, 10.0, 0.2, 0.2, 100, 2, 1, native.stream_ptr()), 'ctl_init')
    off = L['h32']
    ctl[off:off + 4] = torch.frombuffer(bytearray(np.float32(0.1).tobytes()), dtype=torch.uint8).to(device)
    ks = (ctypes.c_void_p * 7)(*[b.data_ptr() for b in bufs[:7]])
    partials = torch.zeros(2 * L['max_seg'] * L['partial_blocks'], dtype=torch.float64, device=device)
    flag = torch.zeros(1, dtype=torch.int32, device=device)
    seg = (native.host_i64([0]), native.host_i64([n]), 1)

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in ev:
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        return statistics.median(a.elapsed_time(b) for a, b in ev) * 1e-3

    t_comb = timed(lambda: lib.node_b200_rk_stage_combine(native.ptr(ctl), native.F32, 5, native.ptr(bufs[8]), native.ptr(bufs[7]),
                                                          ks, 6, n, native.stream_ptr()))
    t_err = timed(lambda: lib.node_b200_rk_error_norm(native.ptr(ctl), native.F32, native.ptr(bufs[7]), native.ptr(bufs[8]), ks, *seg,
                                                      native.ptr(partials), native.ptr(flag), native.stream_ptr()))
    comb_bytes = 7 * n * 4            # y0 + k1,k3..k6 read, y7 written (beta_62 = 0)
    err_bytes = 8 * n * 4             # k1,k3..k7, y0, y1 read; nothing written
    return dict(stage_combine=dict(bound='hbm', achieved=comb_bytes / t_comb / 1e9, peak=pk['hbm'], unit='GB/s',
                                   frac=comb_bytes / t_comb / 1e9 / pk['hbm'], bytes_per_launch=comb_bytes, ms=t_comb * 1e3),
                error_norm=dict(bound='hbm', achieved=err_bytes / t_err / 1e9, peak=pk['hbm'], unit='GB/s',
                                frac=err_bytes / t_err / 1e9 / pk['hbm'], bytes_per_launch=err_bytes, ms=t_err * 1e3),
                peak_source=pk['src'], note='E = 32Mi fp32 per tensor (9 tensors, 1.1 GiB) so every pass streams from HBM')


def train_step_rate(dev, batch, steps, warmup, world=1):
    """cfg3: CIFAR-10 ODENet training step - forward, cross-entropy, odeint_adjoint backward through the native VJP
    kernels, gradient synchronisation (N > 1: adjoint collectives inside the solve + one bucketed all-reduce of the non-ODE
    gradients, node_b200.distributed.sync_gradients), SGD step (reproduce.sh:3-6 hyper-parameters). Every rank runs it on
    its own shard of `batch` images; returns the dict for the JSON line (aggregate images/s, max over ranks)."""
    import torch.distributed as dist
    from node_b200 import solver, caller_grad, distributed as nd
    cg0 = caller_grad.launches
    torch.manual_seed(0)
    net = build_model(dev, adjoint=True, dropout=0.5).train()          # cfg3: reproduce.sh:3-6 (lr 0.1, dropout 0.5, wd 1e-4)
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    rank = int(os.environ.get('RANK', '0'))
    g = torch.Generator().manual_seed(99 + rank)
    x = torch.rand(batch, 3, 32, 32, generator=g).to(dev)
    y = torch.randint(0, 10, (batch,), generator=g).to(dev)
    nfe = [0, 0]
    sent = [0]

    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(net(x), y)
        nfe[0] = net.nfe(reset=True)
        loss.backward()
        nfe[1] = net.nfe(reset=True)
        sent[0] = nd.sync_gradients(net)
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        loss = step()
    b.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tt = torch.tensor([a.elapsed_time(b) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt)
    return dict(images_per_s=world * batch * steps / dt, ms_per_step=1e3 * dt / steps, per_gpu_batch=batch, global_batch=world * batch,
                n_gpus=world, steps=steps, nfe_forward=nfe[0], nfe_backward=nfe[1], adjoint_vjp=solver.last_stats.get('adjoint_vjp'),
                loss=float(loss.detach()), non_ode_gradient_floats_allreduced=sent[0],
                callers_backward='native' if caller_grad.launches > cg0 else 'pytorch',
                note='forward + CE loss + odeint_adjoint backward (native VJP kernels, tol 1e-3) + gradient sync + SGD step; '
                     'downsampler forward AND backward on this repo\'s kernels (node_b200.caller_grad: GroupNorm/ReLU backward, tcgen05 '
                     'data / weight gradients, fused stem backward), pooling / dropout / linear / loss in PyTorch; weak scaling '
                     '(per-GPU batch fixed), device time, max over ranks')


def strong_scaling(net, dev, world, rank):
    """the forward at FIXED global batches 128 / 1024 / 8192 split over the GPUs of the run (batch-global error norm
    all-reduced every attempted step). Device time, max over ranks."""
    import torch.distributed as dist
    out = {}
    for gb in (128, 1024, 8192):
        if gb % world:
            continue
        xb = torch.rand(gb // world, 3, 32, 32, generator=torch.Generator().manual_seed(500 + rank)).to(dev)
        with torch.no_grad():
            for _ in range(3):
                net(xb)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            a.record()
            for _ in range(reps):
                net(xb)
            b.record()
            torch.cuda.synchronize()
        tt = torch.tensor([a.elapsed_time(b) * 1e-3 / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        out['global_batch_%d' % gb] = dict(ms_per_forward=1e3 * float(tt), images_per_s=gb / float(tt), per_gpu_batch=gb // world)
    return out


def incumbent(dev, batch):
    """The reference's eager CUDA path on this same B200: the reference's op sequence - ATen / cuDNN
    fp32 convolutions (allow_tf32 off), native_group_norm, the Python dopri5 loop with its host syncs - with none of this
    repo's kernels (caller kernels off, solver = the oracle's torch restatement on cuda tensors)."""
    os.environ['NODE_B200_CALLERS'] = '0'
    try:
        net = reference_model(dev)
        out = {}
        for b in (128, batch):
            x = torch.rand(b, 3, 32, 32, device=dev)
            with torch.no_grad():
                t = timed_host(lambda: net(x), 2, 5)
            out['batch_%d' % b] = dict(images_per_s=b / t, ms_per_forward=1e3 * t)
        out['note'] = 'eager PyTorch on cuda: cuDNN fp32 convolutions, ATen elementwise RK stages, >= 9 host syncs per attempted step'
        return out
    finally:
        os.environ.pop('NODE_B200_CALLERS', None)


def cfg1_and_cfg3_baselines(dev, threads):
    """BASELINE.json configs[0] (MNIST ODENet inference, batch 128, defined as a CPU configuration) and configs[2] (CIFAR training
    step with the adjoint, batch 128) on the host CPU through the reference's op sequence, beside this repo's GPU path."""
    from node_b200 import models, solver
    torch.set_num_threads(threads)
    out = {}
    x1 = torch.rand(128, 1, 28, 28)
    ref = reference_model('cpu', in_ch=1, downsample='convolution')
    with torch.no_grad():
        t_cpu = timed_host(lambda: ref(x1), 1, 5)
    torch.manual_seed(0)
    net = models.ODENet(1, n_filters=64, downsample='convolution', tol=TOL).eval().to(dev)
    xg = x1.to(dev)
    with torch.no_grad():
        t_gpu = timed_host(lambda: net(xg), 3, 10)
        t_e2e = timed_host(lambda: net(x1.to(dev)).cpu(), 3, 10)
    out['cfg1_mnist_inference_b128'] = dict(cpu_images_per_s=128 / t_cpu, cpu_cores=threads, gpu_images_per_s=128 / t_gpu,
                                            gpu_e2e_images_per_s=128 / t_e2e, nfe=solver.last_stats.get('nfe'), route=solver.last_stats.get('route'),
                                            note='state [128,64,6,6]; wall clock incl. the one status read per solve')
    xb = torch.rand(128, 3, 32, 32)
    yb = torch.randint(0, 10, (128,))
    ref = reference_model('cpu', adjoint=True, train=True)
    opt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)

    def cpu_step():
        opt.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(ref(xb), yb).backward()
        opt.step()

    t_cpu3 = timed_host(cpu_step, 1, 3)
    net = build_model(dev, adjoint=True).train()
    optg = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    xg, yg = xb.to(dev), yb.to(dev)

    def gpu_step():
        optg.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(net(xg), yg).backward()
        optg.step()

    t_gpu3 = timed_host(gpu_step, 2, 5)
    out['cfg3_train_step_b128'] = dict(cpu_images_per_s=128 / t_cpu3, cpu_cores=threads, gpu_images_per_s=128 / t_gpu3,
                                       gpu_ms_per_step=1e3 * t_gpu3, note='the reference default batch (train.py:207), adjoint, SGD')
    return out


def pgd_latency(dev, threads):
    """BASELINE.json configs[4]: the attack loop of adversarial/attack.py:63-72 restated as plain PGD (foolbox is not vendored by
    the reference: SURVEY 8c) - batch 1, 10 iterations of forward + input gradient through the ODE block (odeint_adjoint), signed
    step 0.01, eps 0.03, clip to [0, 1] - at tol 1e-4 .. 1e-1, ms per iteration beside the reference's CPU path."""
    from node_b200 import solver
    torch.set_num_threads(threads)
    x0 = torch.rand(1, 3, 32, 32, generator=torch.Generator().manual_seed(11))
    label = torch.tensor([3])

    def attack(net, device, iters):
        x = x0.to(device).clone()
        lab = label.to(device)
        nfe = 0
        for _ in range(iters):
            x.requires_grad_(True)
            loss = torch.nn.functional.cross_entropy(net(x), lab)
            g, = torch.autograd.grad(loss, x)
            nfe = net.nfe(reset=True)
            x = (x.detach() + 0.01 * g.sign())
            x = torch.min(torch.max(x, x0.to(device) - 0.03), x0.to(device) + 0.03).clamp(0, 1)
        return nfe

    out = {}
    gnet = build_model(dev, adjoint=True)
    cnet = reference_model('cpu', adjoint=True)
    for tol in (1e-4, 1e-3, 1e-2, 1e-1):
        gnet.odeblock.tol = tol
        cnet.odeblock.tol = tol
        nfe_box = [0]
        t_gpu = timed_host(lambda: nfe_box.__setitem__(0, attack(gnet, dev, 10)), 1, 3) / 10
        nfe_gpu = nfe_box[0]
        t_cpu = timed_host(lambda: nfe_box.__setitem__(0, attack(cnet, 'cpu', 3)), 0, 1) / 3
        out['tol_%g' % tol] = dict(gpu_ms_per_iteration=1e3 * t_gpu, cpu_ms_per_iteration=1e3 * t_cpu, nfe_fwd_plus_bwd=nfe_gpu,
                                   nfe_fwd_plus_bwd_cpu=nfe_box[0])
    out['note'] = 'batch 1: replicas only across GPUs (no collective); wall clock per PGD iteration; cpu = %d threads' % threads
    out['adjoint_vjp'] = solver.last_stats.get('adjoint_vjp')
    return out


def other_configs(dev):
    """The other configurations of SURVEY 8(d), measured briefly beside the headline (device-timed, inputs resident):
    cfg4 ICMR feature extraction (10 output times, dense output costs no extra evaluations), the `one-shot`
    downsampler (16x16 maps), and n_filters=256 (paper CIFAR setting: served by the generic route, K2-K6 kernels +
    PyTorch dynamics)."""
    import numpy as np
    from node_b200 import models, solver
    out = {}

    def rate(fn, batch, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return batch * reps / (a.elapsed_time(b) * 1e-3)

    with torch.no_grad():
        torch.manual_seed(0)
        net = models.ODENet(3, n_filters=64, downsample='residual', tol=TOL).eval().to(dev)
        net.to_features_extractor()
        net.odeblock.t1 = np.linspace(0, 1, 10).tolist()
        x = torch.rand(2220, 3, 32, 32, device=dev)
        r = rate(lambda: net(x), 2220)
        out['cfg4_features_t10'] = dict(images_per_s=r, batch=2220, output_times=10, nfe=solver.last_stats.get('nfe'),
                                        route=solver.last_stats.get('route'), features_shape=list(net(x).shape))
        torch.manual_seed(0)
        net = models.ODENet(3, n_filters=64, downsample='one-shot', tol=TOL).eval().to(dev)
        x = torch.rand(1024, 3, 32, 32, device=dev)
        r = rate(lambda: net(x), 1024)
        out['one_shot_16x16'] = dict(images_per_s=r, batch=1024, nfe=solver.last_stats.get('nfe'), route=solver.last_stats.get('route'))
        torch.manual_seed(0)
        net = models.ODENet(3, n_filters=256, downsample='residual', tol=TOL).eval().to(dev)
        x = torch.rand(256, 3, 32, 32, device=dev)
        r = rate(lambda: net(x), 256, reps=3)
        out['n_filters_256'] = dict(images_per_s=r, batch=256, nfe=solver.last_stats.get('nfe'), route=solver.last_stats.get('route'))
        # the paper's CIFAR setting (reproduce.sh:21) at a batch of whole rounds of the wide8 convolution kernel (148 CTAs x 4 images x 4)
        x = torch.rand(2368, 3, 32, 32, device=dev)
        r = rate(lambda: net(x), 2368, reps=3)
        h0 = net.downsample(x)
        rb = rate(lambda: net.odeblock(h0), 2368, reps=3)
        nfe = solver.last_stats.get('nfe')
        out['n_filters_256_b2368'] = dict(images_per_s=r, batch=2368, nfe=nfe, route=solver.last_stats.get('route'), odeblock_ms=2368e3 / rb,
                                          odeblock_algorithmic_tflops=2368 * nfe * 2 * 2 * 9 * 256 * 256 * 64 / (2368.0 / rb) / 1e12,
                                          note='dynamics: wide8 engine (one tcgen05 implicit GEMM per convolution, f16x3); RK stages / error norm: generic route kernels')
        os.environ['NODE_B200_WIDE'] = '0'
        try:
            r0 = rate(lambda: net(x), 2368, reps=2)
            out['n_filters_256_b2368']['cudnn_dynamics_images_per_s'] = r0
        finally:
            os.environ.pop('NODE_B200_WIDE', None)
        del x, h0
        # the paper's CIFAR TRAINING setting (reproduce.sh:21-25: --filters 256 --adjoint --batch-size 128 --dropout 0.5): one SGD step,
        # the ODE block's adjoint on the native wide augmented dynamics (node_b200_wide_vjp) vs the autograd / cuDNN route
        try:
            def wide_step(netw, optw, xw, yw):
                optw.zero_grad(set_to_none=True)
                with torch.enable_grad():
                    torch.nn.functional.cross_entropy(netw(xw), yw).backward()
                optw.step()
            res = {}
            for bt in (128, 1024):
                for mode in ('1', '0'):
                    os.environ['NODE_B200_NATIVE_VJP'] = mode
                    torch.manual_seed(0)
                    netw = models.ODENet(3, n_filters=256, downsample='residual', tol=TOL, adjoint=True, dropout=0.5).train().to(dev)
                    optw = torch.optim.SGD(netw.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
                    xw = torch.rand(bt, 3, 32, 32, device=dev)
                    yw = torch.randint(0, 10, (bt,), device=dev)
                    for _ in range(2):
                        wide_step(netw, optw, xw, yw)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for _ in range(3):
                        wide_step(netw, optw, xw, yw)
                    torch.cuda.synchronize()
                    res['b%d_%s' % (bt, 'native' if mode == '1' else 'autograd_cudnn')] = dict(
                        images_per_s=3 * bt / (time.perf_counter() - t0), adjoint_vjp=solver.last_stats.get('adjoint_vjp'))
                    del netw, optw, xw, yw
            res['note'] = ('whole training step at 256 filters; ODE block forward (wide8) and adjoint (node_b200_wide_vjp) on this '
                           "repo's kernels, the 256-filter downsampler / classifier are PyTorch (the caller kernels serve 64 filters)")
            out['n_filters_256_train_step'] = res
        except Exception as exc:
            out['n_filters_256_train_step'] = dict(error=repr(exc))
        finally:
            os.environ.pop('NODE_B200_NATIVE_VJP', None)
        # training WITHOUT --adjoint (train.py:221 default): gradients through the unrolled solver (node_b200.unrolled:
        # recorded solver loop, native dynamics + VJP kernels) at the reference's batch of 128
        try:
            torch.manual_seed(0)
            netu = models.ODENet(3, n_filters=64, downsample='residual', tol=TOL, adjoint=False, dropout=0.5).train().to(dev)
            optu = torch.optim.SGD(netu.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
            xu = torch.rand(128, 3, 32, 32, device=dev)
            yu = torch.randint(0, 10, (128,), device=dev)

            def ustep():
                optu.zero_grad(set_to_none=True)
                with torch.enable_grad():
                    torch.nn.functional.cross_entropy(netu(xu), yu).backward()
                optu.step()
            for _ in range(2):
                ustep()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                ustep()
            torch.cuda.synchronize()
            out['unrolled_train_step_b128'] = dict(images_per_s=5 * 128 / (time.perf_counter() - t0), route=solver.last_stats.get('route'),
                                                   note='odeint (non-adjoint) under autograd: the reference\'s unrolled gradient, controller terms included')
            del netu, optu
        except Exception as exc:
            out['unrolled_train_step_b128'] = dict(error=repr(exc))
        # batched independent solvers (evaluate.py:109-126 runs batch_size = 1 to record the NFE of every image)
        try:
            from node_b200 import each
            net64 = models.ODENet(3, n_filters=64, downsample='residual', tol=TOL).eval().to(dev)
            h = net64.downsample(torch.rand(512, 3, 32, 32, device=dev))
            tt = torch.tensor([0.0, 1.0], device=dev)
            fn = lambda: each.odeint_each(net64.odeblock.odefunc, h, tt, rtol=TOL, atol=TOL, lanes=32)
            r = rate(fn, 512, reps=2)
            stats = fn()[1]
            import torchdiffeq as _td
            t0 = time.perf_counter()
            for i in range(64):
                _td.odeint(net64.odeblock.odefunc, h[i:i + 1], tt, rtol=TOL, atol=TOL, method='dopri5')
            torch.cuda.synchronize()
            seq = 64 / (time.perf_counter() - t0)
            out['per_sample_solves'] = dict(solves_per_s=r, sequential_batch1_solves_per_s=seq, samples=512, lanes=32,
                                            nfe_min=min(s['nfe'] for s in stats), nfe_max=max(s['nfe'] for s in stats),
                                            note='N independent dopri5 solves (per-sample error norm and step sequence), 32 in flight on private workspaces / streams / captured launch sequences; each equals the batch-1 solve bit for bit')
        except Exception as exc:                      # a bench extra must not take the headline down
            out['per_sample_solves'] = dict(error=repr(exc))
        # retrieval scoring of the 10,000-image test set (evaluate.py:326,339) on a device-resident feature plane
        from node_b200 import retrieval
        feats = torch.rand(10000, 64, device=dev)
        t_cpu = None
        fn = lambda: retrieval.retrieval_scores(retrieval.normalize_features(feats)[0], retrieval.normalize_features(feats)[0])
        r = rate(fn, 1, reps=5)
        f_np = feats.cpu().numpy()
        t0 = time.perf_counter()
        q = f_np / (np.linalg.norm(f_np, axis=-2, keepdims=True) + 1e-7)
        q.dot(q.T)
        t_cpu = time.perf_counter() - t0
        out['retrieval_scores_10k'] = dict(gpu_ms=1e3 / r, cpu_numpy_ms=1e3 * t_cpu, scores_bytes=4 * 10000 * 10000,
                                           write_gb_per_s=4e8 / (1.0 / r) / 1e9,
                                           note='normalise (twice) + 10000x10000x64 fp32 scores; bound by the 400 MB write')
    return out


def small_batch_latency(net, dev, batch=128, reps=30):
    """ODE block at the reference's default batch (train.py:207), where a solve is bound by its ~35 launches:
    direct enqueue against CUDA-graph replay of the same launch sequence (north_star 3)."""
    from node_b200 import solver
    out = dict(batch=batch)
    with torch.no_grad():
        h0 = net.downsample(torch.rand(batch, 3, 32, 32, device=dev))
        for mode in ('0', '1'):
            os.environ['NODE_B200_GRAPH'] = mode
            for _ in range(3):
                net.odeblock(h0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                net.odeblock(h0)
            torch.cuda.synchronize()
            out['direct_ms' if mode == '0' else 'graph_ms'] = 1e3 * (time.perf_counter() - t0) / reps
    os.environ.pop('NODE_B200_GRAPH', None)
    out['launches_per_solve'] = solver.last_stats.get('launches')
    out['note'] = 'wall clock per ODE-block forward incl. the one status read-back per solve'
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--batch', type=int, default=int(os.environ.get('NODE_B200_BENCH_BATCH', 9472)),
                    help='per-GPU batch; 9472 = 8 rounds of 148 SMs x 2 virtual slots x 4 images per super-tile (no tail round; 4736 = 4 rounds: -3 %% throughput)')
    ap.add_argument('--cpu-sample', type=int, default=0, help='images per step of the reference arm / cpu_baseline (0 = the whole per-GPU batch)')
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--skip-cpu', action='store_true')
    ap.add_argument('--quick', action='store_true', help='tuning runs: only the forward, e2e, ODE-block and step-kernel legs')
    ap.add_argument('--train-batch', type=int, default=4736, help='batch of the adjoint training-step measurement (0 = skip)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'native' else args.warmup
    if args.impl == 'reference':
        return run_reference(args)

    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    assert world == args.gpus, 'launch with torchrun --nproc-per-node %d for --gpus %d' % (args.gpus, args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    import __graft_entry__ as entry
    entry.build()
    from node_b200 import solver, distributed as nd
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
        nd.enable()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    pk = peaks()

    net = build_model(dev)
    B = args.batch
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.rand(B, 3, 32, 32, generator=g).pin_memory()
    x_dev = x_host.to(dev)
    logits_host = torch.empty(B, 10).pin_memory()
    solver.PROFILE_STEP_EVENTS = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(step_fn, k):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            step_fn()
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b) * 1e-3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    launches = [0]

    from node_b200 import caller_ops

    def step_resident():
        c0 = caller_ops.launches
        with torch.no_grad():
            net(x_dev)
        launches[0] += solver.last_stats.get('launches', 0) + (caller_ops.launches - c0)

    # End to end: every step copies ITS input batch from pinned host memory and reads ITS logits back to the host. The
    # input of step i+1 is prefetched on a copy stream while step i computes (the usual pinned-memory data-loader pattern:
    # two device buffers, events in both directions); the result read-back is synchronous, once per step.
    copy_stream = torch.cuda.Stream(device=dev)
    xbuf = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_i = [0]

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])          # the forward that last read this buffer is done
            xbuf[slot].copy_(x_host, non_blocking=True)
            copied[slot].record(copy_stream)

    for ev in consumed:
        ev.record()
    prefetch(0)

    def step_e2e():
        cur = e2e_i[0] & 1
        e2e_i[0] += 1
        main = torch.cuda.current_stream()
        main.wait_event(copied[cur])
        prefetch(cur ^ 1)                                    # next step's input travels while this step computes
        with torch.no_grad():
            out = net(xbuf[cur])
        consumed[cur].record(main)
        logits_host.copy_(out, non_blocking=True)
        main.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        step_resident()
    solver.PROFILE_STEP_EVENTS = []
    launches[0] = 0
    if sampler:
        torch.cuda.synchronize()
        sampler.mark()
    t_res = timed_loop(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None
    step_events = solver.PROFILE_STEP_EVENTS
    solver.PROFILE_STEP_EVENTS = None
    n_launch = launches[0]
    stats = dict(solver.last_stats)
    for _ in range(2):
        step_e2e()
    t_e2e = timed_loop(step_e2e, args.steps)

    # ODE block alone (the hot path proper), state resident
    with torch.no_grad():
        h0 = net.downsample(x_dev)
    def step_ode():
        with torch.no_grad():
            net.odeblock(h0)

    t_ode = timed_loop(step_ode, args.steps)

    # legs every rank takes part in (collectives inside): cfg3 training step and the fixed-global-batch forwards
    strong = strong_scaling(net, dev, world, rank) if not args.quick else None
    train = train_step_rate(dev, args.train_batch, max(2, args.steps // 2), 2, world) if args.train_batch > 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total = world * B * args.steps
    value = total / t_res
    # dominant kernel: the fused 6-stage step kernel (tensor bound)
    k_ms = [a.elapsed_time(b) for a, b in step_events]
    k_avg = statistics.mean(k_ms) * 1e-3 if k_ms else float('nan')
    flops_per_launch = 6 * B * FLOP_PER_IMG_PER_NFE_PER_PIXEL * 64
    # Denominator: the fp32-contract convolution is judged against the TF32 dense rate = 1/2 of the measured bf16 GEMM rate
    # (MEASURED_PEAKS.json has no TF32 figure). A ~1 ms kernel at an unthrottled 1965 MHz is a BURST measurement; the sustained
    # figure (power-capped seconds-long GEMM loop) is quoted beside it.
    tf32_peak = pk['bf16'] / 2
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(tpath):
        tj = json.load(open(tpath))                       # one ncu --set full capture; scaled to this run's batch per launch
        traffic = tj.get('k_step_dram_bytes_per_launch')
        if tj.get('k_step_dram_bytes_per_image') and tj.get('batch_of_capture') != B:
            traffic = int(tj['k_step_dram_bytes_per_image'] * B)
    roof = dict(bound='tensor', kernel='k_step8 (dense 8x8 tiling on CTA pairs, tcgen05.mma.cta_group::2; 6 dopri5 stages = 12 '
                                       'implicit-GEMM convs per launch)',
                achieved=flops_per_launch / k_avg / 1e12, peak=tf32_peak, unit='TFLOP/s',
                frac=flops_per_launch / k_avg / 1e12 / tf32_peak, traffic=traffic,
                frac_of_sustained=flops_per_launch / k_avg / 1e12 / (pk['bf16_sus'] / 2),
                frac_of_issued_f16=3 * flops_per_launch / k_avg / 1e12 / pk['bf16'],
                flops_per_launch=flops_per_launch, launch_ms=k_avg * 1e3, launches_timed=len(k_ms),
                share_of_step=sum(k_ms) * 1e-3 / t_res,
                peak_note='TF32 dense peak taken as 1/2 of the %s TFLOP/s burst bf16 GEMM rate (%s; sustained %s); algorithmic FLOPs '
                          'counted once although the fp16 operand split issues 3 products (frac_of_issued_f16 counts them '
                          'against the bf16 rate)' % (pk['bf16'], pk['src'], pk['bf16_sus']))
    line = dict(metric='CIFAR-10 ODENet dopri5 forward throughput', value=value, unit='images/s', n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * t_res / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic', config=workload_config(args),
                e2e=dict(value=total / t_e2e, unit='images/s', h2d_bytes_per_step=x_host.numel() * 4,
                         d2h_bytes_per_step=logits_host.numel() * 4),
                gpu_launches=n_launch, clocks=clocks, roofline=roof,
                odeblock=dict(images_per_s=total / t_ode, ms_per_step=1e3 * t_ode / args.steps, nfe=stats.get('nfe'),
                              n_accept=stats.get('n_accept'), n_reject=stats.get('n_reject')))
    line['strong_scaling'] = strong
    if train is not None:
        line['train_step'] = train
    if world == 1 and not args.quick:
        line['roofline_rk'] = rk_roofline(dev, pk)
        line['latency_b128'] = small_batch_latency(net, dev)
        line['other_configs'] = other_configs(dev)
        line['incumbent_eager_cuda'] = incumbent(dev, B)
        if not args.skip_cpu:
            threads = os.cpu_count() or 1
            sample = B if args.cpu_sample <= 0 else min(B, args.cpu_sample)
            rate, times = cpu_forward_rate(sample, args.cpu_seconds, threads)
            line['cpu_baseline'] = dict(value=rate, unit='images/s', cores=threads, kind='port',
                                        sample='%d-image batches (the per-GPU batch), %d forwards (median), torch-CPU restatement of the '
                                               'reference solver + dynamics (oracle/), same seeds' % (sample, len(times)))
            line['cpu_baselines_other_configs'] = cfg1_and_cfg3_baselines(dev, threads)
            line['cfg5_pgd'] = pgd_latency(dev, threads)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
