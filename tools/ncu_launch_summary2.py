#!/usr/bin/env python
"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel name count / total / mean / share.
   python tools/ncu_launch_summary2.py FILE.csv [--seq]   (--seq: print every launch in order)"""
import collections, csv, re, sys
f = sys.argv[1]
lines = [l for l in open(f) if not l.startswith("==")]
agg = collections.OrderedDict()
tot = 0.0
seq = []
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    grid = r.get("Grid Size", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v; tot += v
    seq.append((name, grid, v))
print(f"{f}: {sum(a[0] for a in agg.values())} launches, {tot:.0f} us")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k[:70]:70s} n={n:5d} total={t:10.1f} us mean={t / n:9.2f} us share={t / tot:.3f}")
if "--seq" in sys.argv:
    for name, grid, v in seq:
        print(f"    {name[:50]:50s} {grid:>18s} {v:9.1f} us")
