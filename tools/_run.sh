mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err ) 2>&1 | grep real; tail -2 gpurun_out/t_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t_bench.json').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f ms %.3f launch_ms %.4f frac %.4f ode_ms %.3f' % (d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['launch_ms'],d['roofline']['frac'],d['odeblock']['ms_per_step']))
print(d['odeblock']); print(d['train_step']); print(d['cpu_baseline']); print(d['clocks']); print(d['latency_b128'])
PY
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/t_ref.json 2> gpurun_out/t_ref.err ) 2>&1 | grep real; cat gpurun_out/t_ref.json | cut -c1-600
