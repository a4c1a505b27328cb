"""Print an ncu `--metrics gpu__time_duration.sum --csv` launch list as a table + per-kernel totals."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); gi = h.index('Grid Size')
tot = 0; agg = collections.OrderedDict(); lst = []
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    n = r[ki]; n = n[n.find('::') + 2:] if '::' in n else n
    n = n.split('(')[0][:40]
    lst.append((r[0], n, r[gi], v / 1e3)); agg[n] = agg.get(n, 0) + v / 1e3; tot += v / 1e3
if '-v' in sys.argv:
    for x in lst: print("%3s %-42s grid=%-16s %8.1f us" % x)
print("total %.1f us over %d launches" % (tot, len(lst)))
for k, v in sorted(agg.items(), key=lambda x: -x[1]): print("%10.1f us %5.1f%% %s" % (v, 100 * v / tot, k))
