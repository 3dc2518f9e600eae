"""Samples per device function (by source line range) from an ncu cuda,sass source export.
python tools/ncu_src_by_func.py export.csv path/to/file.cuh"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
src = open(sys.argv[2]).read().split('\n')
fname = sys.argv[2].split('/')[-1]
# function starts: lines beginning with __device__/__global__/template at col 0
starts = []
for i, l in enumerate(src, 1):
    m = re.match(r'^(?:__device__|__global__|inline|template)', l)
    if m:
        # find name on this or following lines
        txt = ' '.join(src[i - 1:i + 3])
        n = re.search(r'(\w+)\s*\(', txt.split('__forceinline__')[-1].split('__launch_bounds__')[-1].split(')')[-0] if False else txt)
        names = re.findall(r'(\w+)\(', txt)
        names = [x for x in names if x not in ('__launch_bounds__', 'template')]
        starts.append((i, names[0] if names else '?'))
def func_of(line):
    cur = '?'
    for s, n in starts:
        if s <= line: cur = n
        else: break
    return cur
cur = None; agg = {}; tot = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if len(r) > 7 and r[0].isdigit():
        s = int(r[4] or 0)
        key = func_of(int(r[0])) if cur == fname else cur
        agg[key] = agg.get(key, 0) + s; tot += s
for k, v in sorted(agg.items(), key=lambda x: -x[1]):
    print(f"{v / tot * 100:5.1f}%  {k}")
