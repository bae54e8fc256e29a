#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/wide_one.py <<'PY'
import os, sys
sys.path.insert(0, 'neural-ode-features_b200'); sys.path.insert(0, '.')
import torch
import __graft_entry__ as e; e.build()
from node_b200 import models, solver
torch.manual_seed(0)
net = models.ODENet(3, n_filters=256, downsample='residual', tol=1e-3).eval().cuda()
x = torch.rand(2048, 3, 32, 32, device='cuda')
with torch.no_grad():
    h0 = net.downsample(x)
    for _ in range(2): net.odeblock(h0)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    net.odeblock(h0)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02ab_wide_launches.csv python /tmp/wide_one.py > gpurun_out/r02ab_wide.log 2>&1
python tools/launch_agg.py gpurun_out/r02ab_wide_launches.csv 20
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_wide_conv -s 4 -c 1 -f -o gpurun_out/r02ab_k_wide_conv python /tmp/wide_one.py >> gpurun_out/r02ab_wide.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_wide_gn_operand -s 4 -c 1 -f -o gpurun_out/r02ab_k_wide_gn python /tmp/wide_one.py >> gpurun_out/r02ab_wide.log 2>&1
tail -3 gpurun_out/r02ab_wide.log
