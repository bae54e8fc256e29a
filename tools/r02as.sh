#!/bin/bash
python -c "import sys; sys.path.insert(0,'.'); import __graft_entry__ as g; print('stale', g._stale())"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
BATCHES=1,8,128,160,300 timeout 200 python tools/vjp_latency.py 2>&1 | tail -5
NODE_B200_STRIP_GS=0 BATCHES=8,128,160 timeout 200 python tools/vjp_latency.py 2>&1 | tail -3
timeout 300 python bench.py --quick 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', d['value'], 'e2e', d['e2e']['value'], 'lat', d.get('latency_b128'))
"
