mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/t_pytest.log
timeout 600 python bench.py --skip-cpu > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'launch_ms',d['roofline']['launch_ms'],'frac',d['roofline']['frac'],'ode',d['odeblock'])
PY
