"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count, total time and share.
    python tests/summarize_launches.py gpurun_out/launches.csv profiles/r02_launch_list_summary.json"""
import csv
import json
import re
import sys
from collections import OrderedDict

src, dst = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
rows = []
with open(src) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = r['Kernel Name']
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'\(.*$', '', name)
    name = name.replace('vp::', '')
    rows.append((name, float(r['Metric Value']) / 1e6, r['Grid Size'], r['Block Size']))
total = sum(t for _, t, _, _ in rows)
by = OrderedDict()
for name, t, grid, blk in rows:
    d = by.setdefault(name, dict(n=0, ms=0.0))
    d['n'] += 1
    d['ms'] += t
out = dict(source=src, launches=len(rows), sum_ms=total,
           by_kernel=OrderedDict((k, dict(n=v['n'], ms=round(v['ms'], 4), share=round(v['ms'] / total, 4)))
                                 for k, v in sorted(by.items(), key=lambda kv: -kv[1]['ms'])))
top = sorted(rows, key=lambda r: -r[1])[:25]
out['top_launches'] = [dict(kernel=n, ms=round(t, 4), grid=g, block=b) for n, t, g, b in top]
if dst:
    json.dump(out, open(dst, 'w'), indent=1)
print('launches %d, sum %.2f ms' % (len(rows), total))
for k, v in list(out['by_kernel'].items())[:24]:
    print('%6.2f%% %8.3f ms %5d  %s' % (100 * v['share'], v['ms'], v['n'], k))
