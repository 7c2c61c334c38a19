#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
    a = agg.setdefault(name, [0, 0.0, row['Grid Size'], row['Block Size']]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print('kernel,launches,total_us,avg_us,share,grid,block')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('"%s",%d,%.1f,%.2f,%.4f,"%s","%s"' % (k, a[0], a[1], a[1] / a[0], a[1] / tot, a[2], a[3]))
