#!/bin/bash
show() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value %.0f img/s  k_step %.3f ms x%d  frac %.3f  odeblock %.0f img/s %.2f ms' % (d['value'], r['launch_ms'], r['launches_timed'], r['frac'], d['odeblock']['images_per_s'], d['odeblock']['ms_per_step']))
"; }
for b in 592 1184 2368; do
  echo "== batch $b"; timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu --train-batch 0 --batch $b 2>/dev/null | show
done
for lag in 1 3; do
  echo "== lag $lag"; NODE_B200_STEP8_LAG=$lag timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu --train-batch 0 --batch 4736 2>/dev/null | show
done
