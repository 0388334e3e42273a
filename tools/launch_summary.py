"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one training step, per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]
kn, mv, gs = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size')
seq = [(r[kn].split('(')[0].replace('void ', '')[:34], r[gs], float(r[mv]) / 1e3) for r in rows[hi + 1:] if len(r) > mv]
idx = [i for i, s in enumerate(seq) if s[0].startswith('stem_fwd')]
one = seq[idx[0]:idx[1]] if len(idx) > 1 else seq
agg = collections.defaultdict(lambda: [0, 0.0])
for k, g, t in one:
    agg[k][0] += 1
    agg[k][1] += t
tot = sum(t for _, _, t in one)
print("one step: %d launches, %.1f us of kernel time (cold-cache, serialised)" % (len(one), tot))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-36s %4d %9.1f us  %5.1f%%  avg %7.2f" % (k, c, t, 100 * t / tot, t / c))
if len(sys.argv) > 2:
    for i, s in enumerate(one):
        print(i, "%-34s %-14s %7.1f" % s)
