"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: python tools/launch_agg.py file.csv [top]"""
import collections
import csv
import re
import sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if hdr is None:
        if 'Kernel Name' in r:
            hdr = r
        continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(d['Metric Value'].replace(',', ''))
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}[d['Metric Unit']]
    name = re.sub(r'\(.*', '', d['Kernel Name'])[:70]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print('total %.3f ms in %d launches' % (tot, sum(v[0] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%-72s %4d %8.3f ms %5.1f%%  (%.3f each)' % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
