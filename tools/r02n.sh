#!/bin/bash
# A/B of the L2 eviction hints and the slot lag of k_step8 at the bench batch (results: stdout)
mkdir -p gpurun_out
show() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value %.0f img/s  e2e %.0f  k_step %.3f ms x%d  frac %.3f  odeblock %.0f img/s %.2f ms' % (d['value'], d['e2e']['value'], r['launch_ms'], r['launches_timed'], r['frac'], d['odeblock']['images_per_s'], d['odeblock']['ms_per_step']))
"; }
for cfg in "1 0" "0 0" "1 3" "0 3" "1 1" "1 0" "0 0"; do
  set -- $cfg
  echo "== L2 hints $1 lag $2"; NODE_B200_STEP8_L2HINT=$1 NODE_B200_STEP8_LAG=$2 timeout 300 python bench.py --steps 10 --warmup 3 --quick --train-batch 0 --batch 4736 2>/dev/null | show
done 2>&1 | tee gpurun_out/r02n_hints.txt
