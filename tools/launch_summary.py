#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.
    python tools/launch_summary.py profiles/x_launches.csv [--seq FIRST LAST]"""
import csv, re, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
if '--seq' in sys.argv:
    i = sys.argv.index('--seq')
    for r in rows[int(sys.argv[i + 1]):int(sys.argv[i + 2])]:
        print(f"{r[0]:>5s} {re.sub(r'[(<].*', '', r[4])[:44]:44s} grid={r[8]:16s} {int(r[-1]) / 1e3:9.1f} us")
    sys.exit(0)
agg = {}
for r in rows:
    n = re.sub(r'[(<].*', '', r[4])[:50]
    a = agg.setdefault(n, [0, 0])
    a[0] += 1
    a[1] += int(r[-1])
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot / 1e3:.1f} us total")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{k:52s} n={v[0]:4d} {v[1] / 1e3:10.1f} us {100 * v[1] / tot:5.1f}%  avg {v[1] / v[0] / 1e3:8.1f} us")
