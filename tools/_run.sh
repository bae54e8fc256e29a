mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "vjp or adjoint" > gpurun_out/t_pytest2.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/t_pytest2.log
python - <<'PY'
import sys, os
sys.path[:0]=[os.path.join(os.getcwd(),'neural-ode-features_b200'), os.getcwd()]
import torch
import __graft_entry__ as e; e.build()
import bench
torch.backends.cudnn.allow_tf32=False
r = bench.train_step_rate(torch.device('cuda',0), 4440, 5, 3)
print('train %.1f img/s  %.2f ms nfe_b %d' % (r['images_per_s'], r['ms_per_step'], r['nfe_backward']))
PY
