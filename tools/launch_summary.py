#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid).
usage: python tools/launch_summary.py launches.csv [first_fraction_to_skip=0.5] [top=30]"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr, data = rows[0], rows[1:]
iN, iV, iU, iG = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Metric Unit', 'Grid Size'))
skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
data = data[int(len(data) * skip):]
def us(r):
    v = float(r[iV].replace(',', '')); u = r[iU]
    return v / 1000 if u.startswith('n') else v * 1000 if u.startswith('m') else v
agg, cnt = collections.Counter(), collections.Counter()
for r in data:
    k = re.sub(r'\(.*', '', r[iN]).replace('void ', '').replace('<unnamed>::', '')[:48]
    agg[(k, r[iG])] += us(r); cnt[(k, r[iG])] += 1
print(f"{len(data)} launches, {sum(agg.values()):.1f} us")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f"{v:9.1f} us  n={cnt[key]:3d}  avg {v / cnt[key]:8.1f}  {key[0]} {key[1]}")
