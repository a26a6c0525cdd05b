"""Mean duration per kernel name from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]
kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) != len(hdr):
        continue
    v = float(r[mv].replace(',', ''))
    if r[mu] in ('ns', 'nsecond'):
        v /= 1e3
    elif r[mu] in ('ms', 'msecond'):
        v *= 1e3
    agg.setdefault(r[kn].split('(')[0], []).append(v)
for k, v in agg.items():
    print(f'{k[-60:]:60s} n={len(v):3d} mean {sum(v)/len(v):9.1f} us  last {v[-1]:9.1f} us')
