mkdir -p gpurun_out
( time timeout 600 python bench.py --skip-cpu --train-batch 0 > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err ) 2>&1 | grep real; echo "bench exit $?"; tail -5 gpurun_out/t_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t_bench.json').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f ms %.3f' % (d['value'],d['e2e']['value'],d['ms_per_step']))
for k,v in d['other_configs'].items(): print(k, v)
PY
