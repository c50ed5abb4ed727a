"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/*_launches_summary*.txt)."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]
ni, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[ui], 1e-6)
    name = r[ni].split('(')[0].replace('void ', '')
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print('%-76s %6s %12s %7s' % ('kernel', 'count', 'total ms', 'share'))
for k, v in tot.most_common(30):
    print('%-76s %6d %12.3f %6.1f%%' % (k[:76], cnt[k], v, 100 * v / total))
print('total %.3f ms' % total)
