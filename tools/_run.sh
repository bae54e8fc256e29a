mkdir -p gpurun_out
timeout 60 tools/issue_bench 2>&1 | grep "grid 148"
timeout 900 python -m pytest tests -m gpu -x -q -k "fused or vjp or callers" > gpurun_out/t_pytest2.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/t_pytest2.log
timeout 600 python bench.py --skip-cpu --train-batch 0 > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; echo "bench exit $?"; tail -3 gpurun_out/t_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t_bench.json').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f ms %.3f launch_ms %.4f frac %.4f ode_ms %.3f launches %d' % (d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['launch_ms'],d['roofline']['frac'],d['odeblock']['ms_per_step'],d['gpu_launches']))
PY
