#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_rejects.py -x -q 2>&1 | tail -3
show() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value %.0f img/s  e2e %.0f  k_step %.3f ms x%d  frac %.3f  odeblock %.0f img/s %.2f ms' % (d['value'], d['e2e']['value'], r['launch_ms'], r['launches_timed'], r['frac'], d['odeblock']['images_per_s'], d['odeblock']['ms_per_step']))
"; }
for i in 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --quick --train-batch 0 --batch 4736 2>/dev/null | show
done
