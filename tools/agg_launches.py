"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name.
python tools/agg_launches.py file.csv [steps]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
h = rows[hdr]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
agg, tot = {}, 0.0
for r in rows[hdr + 2:]:
    if len(r) <= vi:
        continue
    n = re.sub(r'\(.*', '', r[ki])[:80]
    v = float(r[vi].replace(',', ''))
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
for n, a in sorted(agg.items(), key=lambda x: -x[1][1])[:28]:
    print(f"{a[1] / 1e3 / steps:9.1f} us/step {a[0] / steps:6.1f} launches {a[1] / tot * 100:5.1f}%  {n}")
print(f"total {tot / 1e3 / steps:.1f} us/step, {sum(a[0] for a in agg.values()) / steps:.0f} launches/step")
