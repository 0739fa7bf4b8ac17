"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; data = rows[hi + 1:]
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi: continue
    n = r[ki][:60]; v = float(r[vi].replace(',', ''))
    if r[ui] == 'ns': v /= 1000
    elif r[ui] == 'ms': v *= 1000
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for n, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print("%-62s n=%4d total=%10.1f us avg=%9.2f us  %5.1f%%" % (n, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
