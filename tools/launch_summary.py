#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel."""
import collections, csv, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r'\(.*', '', r[ki]).replace('void ', '')
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(r[ui], 1.0)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print('%-58s %6s %12s %7s' % ('kernel', 'count', 'total us', 'share'))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print('%-58s %6d %12.1f %6.1f%%' % (k[:58], a[0], a[1], 100 * a[1] / tot))
print('%-58s %6d %12.1f' % ('total', sum(a[0] for a in agg.values()), tot))
