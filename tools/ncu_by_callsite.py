#!/usr/bin/env python
"""Warp-stall samples of the frame kernel by call site (ncu --set full --import-source on capture joined with nvdisasm -gi of the
same build). Usage: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_by_callsite.py src.csv"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "leaxer-qwen3-tts_b200", "csrc", "liblqt_b200.so")
SRC = os.path.join(ROOT, "leaxer-qwen3-tts_b200", "csrc", "frame_kernel.cuh")


def main():
    src = open(SRC).read().split("\n")
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=td, check=True, stdout=subprocess.DEVNULL)
        cubin = os.path.join(td, [f for f in os.listdir(td) if f.endswith(".cubin")][0])
        txt = subprocess.run(["nvdisasm", "-gi", cubin], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode().split("\n")
    starts = [i for i, l in enumerate(txt) if l.startswith(".text.")]
    s = next(i for i in starts if txt[i].startswith(".text._ZN3lqt12frame_kernelILi3ELi4"))
    e = next(i for i in starts if i > s)
    ct0 = next(i for i, l in enumerate(src, 1) if "LQT_DEVINL void consume_token" in l)
    ct1 = next(i for i, l in enumerate(src, 1) if i > ct0 and l.startswith("}"))
    lab_of, chain, fresh = {}, [], True
    for ln in txt[s:e]:
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            if fresh:
                chain, fresh = [], False
            chain.append((os.path.basename(m.group(1)), int(m.group(2)), os.path.basename(m.group(3)) if m.group(3) else None,
                          int(m.group(4)) if m.group(4) else None))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            fresh = True
            lab = None
            for f, l, inf, inl in chain:
                if inf == "frame_kernel.cuh" and inl and ct0 <= inl <= ct1:
                    lab = ("consume_token", inl)
                    break
            if lab is None and chain:
                f, l, inf, inl = chain[-1]
                lab = ("kernel", inl if inl else l)
            lab_of[int(m.group(1), 16)] = lab
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    base = int(data[0][ix["Address"]], 16)
    cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    agg, ex = collections.defaultdict(collections.Counter), collections.Counter()
    for r in data:
        a = int(r[ix["Address"]], 16) - base
        ex[lab_of.get(a)] += int(r[ix["Instructions Executed"]])
        for c in cols:
            v = int(r[ix[c]])
            if v:
                agg[lab_of.get(a)][c] += v
    allsmp = sum(sum(c.values()) for c in agg.values())
    tot_by = collections.Counter()
    for c in agg.values():
        tot_by.update(c)
    print("all samples by reason:", {k[6:]: round(v / allsmp, 3) for k, v in tot_by.most_common(10)})
    nb = lambda cn: sum(v for k, v in cn.items() if k != "stall_barrier")
    tot = sum(nb(c) for c in agg.values())
    print(f"non-barrier samples {tot} of {allsmp}; share of the non-barrier samples and executed warp-instructions by call site:")
    totex = sum(ex.values())
    for lab, cn in sorted(agg.items(), key=lambda x: -nb(x[1]))[:24]:
        top = ", ".join(f"{k[6:]}={v}" for k, v in cn.most_common(5) if k != "stall_barrier")
        l = lab[1] if lab else 0
        print(f"{nb(cn) / tot * 100:5.1f}% smp {ex[lab] / totex * 100:5.1f}% exec  {(lab[0] if lab else '?'):13s}:{l:5d}  {src[l - 1].strip()[:64] if l else '':64s} | {top}")


if __name__ == "__main__":
    main()
