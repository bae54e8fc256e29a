#!/bin/bash
for b in 4736 9472 14208; do
timeout 300 python bench.py --quick --skip-cpu --batch $b --train-batch 0 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('batch', d['config']['per_gpu_batch'], 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'ms/step %.3f' % d['ms_per_step'], 'step launch %.4f' % d['roofline']['launch_ms'], 'frac %.3f' % d['roofline']['frac'])
"
done
