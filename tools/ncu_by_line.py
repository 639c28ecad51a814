#!/usr/bin/env python
"""Join ncu's per-SASS-instruction samples (ncu -i X.ncu-rep --page source --csv --print-source sass) with nvdisasm -g line
info to get stall samples and executed instructions per CUDA source line / per function region.
Usage: python tools/ncu_by_line.py sass.csv lines.txt [top]"""
import csv, re, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows[2:]:
    try: data.append((int(r[ia], 16), int(r[isamp]), int(r[iex])))
    except Exception: pass
base = data[0][0]
line_of = {}
cur = None
for ln in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", ln)
    if m: line_of[int(m.group(1), 16)] = cur
agg = defaultdict(lambda: [0, 0])
for a, n, e in data:
    k = line_of.get(a - base)
    agg[k][0] += n; agg[k][1] += e
tot = sum(v[0] for v in agg.values()); totex = sum(v[1] for v in agg.values())
src = {}
for f in set(k[0] for k in agg if k):
    try: src[f] = open("/root/repo/leaxer-qwen3-tts_b200/csrc/" + f).read().split("\n")
    except Exception: src[f] = []
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
print(f"total samples {tot}, executed warp-instructions {totex}")
for k, (n, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src.get(k[0], [""] * 99999)[k[1] - 1].strip()[:95] if k and k[1] - 1 < len(src.get(k[0], [])) else ""
    print(f"{(k[0] if k else '?'):18s}:{(k[1] if k else 0):5d} {100*n/tot:5.1f}% smp  {100*e/totex:5.1f}% exec   {text}")
