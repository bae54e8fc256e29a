#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02fin3_bench.json 2> gpurun_out/r02fin3_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02fin3_bench.json') if l.startswith('{')][0])
print('value %.0f e2e %.0f frac %.3f traffic %s train %.0f cfg5 %.2f b128 %.0f lat %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['train_step']['images_per_s'], d['cfg5_pgd']['tol_0.001']['gpu_ms_per_iteration'], d['cpu_baselines_other_configs']['cfg3_train_step_b128']['gpu_images_per_s'], d['latency_b128']['graph_ms']))
print(json.dumps(d['other_configs'].get('unrolled_train_step_b128')), json.dumps(d['other_configs'].get('n_filters_256_train_step'))[:300])
PY
