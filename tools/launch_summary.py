#!/usr/bin/env python3
"""Aggregate an ncu launch list (gpu__time_duration.sum) of tools/step_times.py by kernel for the LAST complete step (dev tool)."""
import csv, re, sys, collections
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = list(csv.reader(lines)); hdr = r[0]
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
data = [(x[ki], float(x[vi].replace(',', ''))) for x in r[1:]]
names = [d[0] for d in data]
idx = [i for i, n in enumerate(names) if 'k_bbox_partial' in n]
step = data[idx[-2]:idx[-1]] if len(idx) >= 2 else data
agg = collections.OrderedDict()
for n, t in step:
    n = re.sub(r'\(.*', '', n).replace('prb::', '').replace('void ', '')
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t
print('total ms', sum(t for _, t in step) / 1e6, 'launches', len(step))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f'{t/1e6:8.3f} ms {c:4d}  {n[:90]}')
