"""Hot source lines of an ncu `--page source --print-source cuda,sass --csv` export.
python tools/ncu_src_hot.py export.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
agg = {}
tot = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 7 and r[0].isdigit():
        num = lambda x: int(x) if x.strip().lstrip('-').isdigit() else 0
        samples = num(r[4]); inst = num(r[7])
        key = (cur, int(r[0]), r[1].strip()[:110])
        a = agg.setdefault(key, [0, 0]); a[0] += samples; a[1] += inst
        tot += samples
for (f, ln, src), (s, i) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{s / tot * 100:5.1f}% {i:>11d} inst  {f}:{ln}  {src}")
print("total samples", tot)
