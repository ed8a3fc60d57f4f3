"""Top SASS lines by warp-stall samples from an `ncu --page source --csv` export (needs -lineinfo / --import-source)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
data = []
for n, r in enumerate(rows[2:]):
    try:
        data.append((int(r[isamp]), int(r[iex]), n, r[isrc].strip()))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print('total samples', tot, 'instructions', len(data))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for s, ex, n, src in sorted(data, reverse=True)[:top]:
    print(f'{100 * s / max(tot, 1):5.1f}%  {s:6d}  ex={ex:8d}  #{n:5d}  {src[:110]}')
