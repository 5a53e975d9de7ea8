"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of one steady-state forward step."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith('==')))
hdr = rows[0]
ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
data = [(r[ki].split('(')[0].replace('stair::', '').replace('void ', '').replace('<unnamed>::', '')[:44], r[gi], float(r[vi].replace(',', '')) / 1e3)
        for r in rows[1:] if len(r) > vi]
idx = [i for i, d in enumerate(data) if 'group_hist' in d[0]]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 2
s, e = idx[which], idx[which + 1]
step = data[s:e]
if len(sys.argv) > 3:
    for d in step:
        print('%-46s grid %-14s %8.1f us' % d)
agg = collections.OrderedDict()
for k, _, v in step:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
print('launches %d  total %.1f us' % (len(step), tot))
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print('%-46s n %3d  %8.1f us  %5.1f%%' % (k, v[0], v[1], 100 * v[1] / tot))
