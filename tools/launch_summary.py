"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (shares of the captured window)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[h]
ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r'\(CUtensorMap.*|\(.*', '', r[ki]).replace('void ', '')
    agg[name][0] += 1
    agg[name][1] += float(r[vi].replace(',', ''))
tot = sum(v[1] for v in agg.values())
print(f'{"us":>10} {"n":>5} {"share":>6}  kernel')
for name, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f'{t / 1e3:10.1f} {c:5d} {100 * t / tot:5.1f}%  {name[:110]}')
print(f'{tot / 1e3:10.1f} total us over {sum(v[0] for v in agg.values())} launches')
