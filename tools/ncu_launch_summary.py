#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv).
Usage: python tools/ncu_launch_summary.py launches.csv  -> markdown table on stdout."""
import collections
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.defaultdict(dict)
    name = {}
    for r in rows:
        per[r[ix["ID"]]][r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
        name[r[ix["ID"]]] = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("lqt::", "").replace("(int)", "")
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for i, m in per.items():
        a = agg[name[i]]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0) / 1e3        # ns -> us
        a[2] += m.get("dram__bytes_read.sum", 0.0)
        a[3] += m.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | share | avg us | dram read MB/launch | dram write MB/launch |")
    print("|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[1] / tot:.3f} | {a[1] / a[0]:.2f} | {a[2] / a[0] / 1e6:.2f} | {a[3] / a[0] / 1e6:.2f} |")


if __name__ == "__main__":
    main()
