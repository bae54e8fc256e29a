#!/bin/bash
# Round-2 quick check on a B200: parity tests of the fused route + the step kernel's launch time for both 8x8 engines.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -8
for b in 4736 4440; do
  echo "== dense engine, batch $b"; timeout 300 python bench.py --steps 5 --warmup 3 --skip-cpu --train-batch 0 --batch $b 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value %.0f img/s  e2e %.0f  k_step %.3f ms x%d  frac %.3f  odeblock %.0f img/s %.2f ms nfe %s' % (d['value'], d['e2e']['value'], r['launch_ms'], r['launches_timed'], r['frac'], d['odeblock']['images_per_s'], d['odeblock']['ms_per_step'], d['odeblock']['nfe']))
    else: print(l.rstrip())
"
done
echo "== strip engine, batch 4440"; NODE_B200_STEP8=0 timeout 300 python bench.py --steps 5 --warmup 3 --skip-cpu --train-batch 0 --batch 4440 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value %.0f img/s  e2e %.0f  k_step %.3f ms x%d  frac %.3f  odeblock %.0f img/s %.2f ms' % (d['value'], d['e2e']['value'], r['launch_ms'], r['launches_timed'], r['frac'], d['odeblock']['images_per_s'], d['odeblock']['ms_per_step']))
    else: print(l.rstrip())
"
