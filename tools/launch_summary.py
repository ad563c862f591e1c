#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and mean time per kernel, share per class."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
per = collections.OrderedDict()
for r in rows:
    name = re.sub(r'\(.*', '', r[4]).replace('void ', '')
    per.setdefault(name, []).append(float(r[-1]) / 1e6)
tot = sum(sum(v) for v in per.values())
print('kernel, launches, total_ms, mean_ms, share_of_listed_time')
cls = collections.Counter()
for k, v in per.items():
    print(f'{k}, {len(v)}, {sum(v):.3f}, {sum(v)/len(v):.3f}, {sum(v)/tot:.3f}')
    c = 'gauss_xy' if 'gauss_xy' in k else 'gauss_z' if 'gauss_z' in k else 'hessian_eigen' if 'hessian' in k else 'j8' if 'j8' in k else 'other'
    cls[c] += sum(v)
print('# per class share: ' + ', '.join(f'{c} {t/tot:.3f}' for c, t in cls.items()))
