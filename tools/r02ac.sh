#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/wide_one.py <<'PY'
import os, sys
sys.path.insert(0, 'neural-ode-features_b200'); sys.path.insert(0, '.')
import torch
import __graft_entry__ as e; e.build()
from node_b200 import models, solver
torch.manual_seed(0)
B = int(os.environ.get('WB', '2368'))
net = models.ODENet(3, n_filters=256, downsample='residual', tol=1e-3).eval().cuda()
x = torch.rand(B, 3, 32, 32, device='cuda')
with torch.no_grad():
    for _ in range(2): net(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    net(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02ac_wide_launches.csv python /tmp/wide_one.py > gpurun_out/r02ac_wide.log 2>&1
python tools/launch_agg.py gpurun_out/r02ac_wide_launches.csv 28
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_wide_gn -s 4 -c 3 -f -o gpurun_out/r02ac_k_wide_gn python /tmp/wide_one.py >> gpurun_out/r02ac_wide.log 2>&1
tail -3 gpurun_out/r02ac_wide.log
