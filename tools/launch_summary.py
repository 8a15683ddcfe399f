"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_summary.py file.csv [skip_fraction]"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[1:]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
data = data[int(len(data) * frac):]
agg = collections.OrderedDict(); tot = 0.0
for r in data:
    k = r[ix['Kernel Name']].split('(')[0].replace('void ', '')[:40]; v = float(r[ix['Metric Value']]); u = r[ix['Metric Unit']]
    v = v / 1e6 if u == 'ns' else (v / 1e3 if u == 'us' else v)
    a = agg.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v); tot += v
for k, a in agg.items():
    print(f"{k:42s} n={a[0]:4d} total {a[1]:9.3f} ms  {100*a[1]/tot:5.1f}%  max {a[2]:.3f} ms")
print(f"total {tot:.3f} ms over {len(data)} launches (serialised, cold-cache: compare shares)")
