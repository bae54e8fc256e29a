"""Per-source-line warp-stall samples from an `ncu --page source --csv --print-source cuda,sass` dump."""
import csv, sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0

path = sys.argv[1]; thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
rows = list(csv.reader(open(path)))
cur_file = None; hdr = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; si = hdr.index('# Samples'); continue
    if hdr is None or len(r) <= si: continue
    if r[0].isdigit():
        key = (cur_file, int(r[0]), r[1])
        a = agg.setdefault(key, [0, {}])
        a[0] += num(r[si])
        for name in ('stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_barrier', 'stall_math', 'stall_mio', 'stall_branch_resolving', 'stall_not_selected', 'stall_selected'):
            j = hdr.index(name)
            a[1][name] = a[1].get(name, 0) + num(r[j])
tot = sum(a[0] for a in agg.values())
print('total samples', tot)
for (f, ln, src), a in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if a[0] >= thresh * tot:
        top = sorted(a[1].items(), key=lambda kv: -kv[1])[:2]
        print('%-18s %4d %6.1f%%  %-28s %s' % (f, ln, 100.0 * a[0] / tot, ' '.join('%s=%d' % (k[6:], v) for k, v in top), src.strip()[:100]))
