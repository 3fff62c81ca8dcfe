"""per-kernel totals of the SECOND half of an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/run_full.py runs every config twice)"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
H = rows[hdr]; k = H.index('Kernel Name'); v = H.index('Metric Value')
data = rows[hdr + 1:]
half = len(data) // 2 if len(sys.argv) < 3 else 0
tot = collections.OrderedDict()
for r in data[half:]:
    n = r[k].split('(')[0][:70]
    tot.setdefault(n, [0, 0.0]); tot[n][0] += 1; tot[n][1] += float(r[v].replace(',', ''))
print('total ms', sum(t for c, t in tot.values()) / 1e6, 'launches', sum(c for c, t in tot.values()))
for n, (c, t) in sorted(tot.items(), key=lambda x: -x[1][1])[:18]:
    print(f"{t/1e6:9.3f} ms x{c:4d}  {n}")
