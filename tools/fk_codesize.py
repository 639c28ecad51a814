#!/usr/bin/env python
"""Static code footprint of the persistent frame kernel by call site (instruction-cache diet aid).
Usage (no GPU needed): python tools/fk_codesize.py   -- needs cuobjdump/nvdisasm and the built liblqt_b200.so (-lineinfo)."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "leaxer-qwen3-tts_b200", "csrc", "liblqt_b200.so")
SRC = os.path.join(ROOT, "leaxer-qwen3-tts_b200", "csrc", "frame_kernel.cuh")


def main():
    kern = sys.argv[1] if len(sys.argv) > 1 else "_ZN3lqt12frame_kernelILi3ELi4"
    src = open(SRC).read().split("\n")
    ct0 = next(i for i, l in enumerate(src, 1) if "LQT_DEVINL void consume_token" in l)
    ct1 = next(i for i, l in enumerate(src, 1) if i > ct0 and l.startswith("}"))
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=td, check=True, stdout=subprocess.DEVNULL)
        cubin = os.path.join(td, [f for f in os.listdir(td) if f.endswith(".cubin")][0])
        txt = subprocess.run(["nvdisasm", "-gi", cubin], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode().split("\n")
    starts = [i for i, l in enumerate(txt) if l.startswith(".text.")]
    s = next(i for i in starts if txt[i].startswith(".text." + kern))
    e = next((i for i in starts if i > s), len(txt))
    chain, fresh, cnt = [], True, collections.Counter()
    for ln in txt[s:e]:
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            if fresh:
                chain, fresh = [], False
            chain.append((os.path.basename(m.group(1)), int(m.group(2)), os.path.basename(m.group(3)) if m.group(3) else None,
                          int(m.group(4)) if m.group(4) else None))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            fresh = True
            if not chain:
                continue
            lab = None
            for f, l, inf, inl in chain:                      # the call site inside consume_token, if any
                if inf == "frame_kernel.cuh" and inl and ct0 <= inl <= ct1:
                    lab = ("consume_token", inl)
                    break
            if lab is None:
                f, l, inf, inl = chain[-1]
                lab = ("kernel", inl if inl else l)
            cnt[lab] += 1
    tot = sum(cnt.values())
    print(f"{kern}: {tot} instructions = {tot * 16 / 1024:.0f} KB")
    for (k, l), n in sorted(cnt.items(), key=lambda x: -x[1])[:32]:
        print(f"{k:14s} line {l:5d} {n:6d} instr {n * 16 / 1024:6.1f} KB   {src[l - 1].strip()[:100]}")


if __name__ == "__main__":
    main()
