#!/usr/bin/env python
"""Phase timeline of one CTA of the persistent frame kernel (csrc/frame_kernel.cuh fk_mark).
Usage (GPU box): python tools/fk_timeline.py [--frames 12] [--cta 0] [--spec 0.6b]
Prints, per (stack, phase kind), the mean SM-clock cycles spent in: wait (grid barrier), stage (input
assembly: loads / norm / attention), gemv (weights from the ring), and the per-frame totals."""
import argparse
import collections
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("LQT_B200_LIB", os.path.join(ROOT, "leaxer-qwen3-tts_b200", "csrc", "liblqt_b200_prof.so"))   # the build with timeline marks
from __graft_entry__ import load_package  # noqa: E402

load_package()
from leaxer_qwen3_tts_b200 import engine, modelspec as ms  # noqa: E402

KINDS = {1: "A qkv", 2: "B attn", 3: "C o-proj", 4: "D gate/up", 5: "E down", 6: "head", 7: "sample", 8: "inproj"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--cta", type=int, default=0)
    ap.add_argument("--spec", default="0.6b")
    ap.add_argument("--mhz", type=float, default=1965.0)
    a = ap.parse_args()
    spec = {"0.6b": ms.spec_0p6b, "1.7b": ms.spec_1p7b, "tiny": ms.spec_tiny}[a.spec](0)
    mdir = ms.generate_model_dir(ms.default_model_dir(spec), spec)
    eng = engine.Engine(mdir)
    ids = engine.wrap_text_ids(ms.synthetic_text_ids(90, 1234))
    prompt, trailing, pad = eng.build_prompt(ids, "en")
    sp = eng.sampling(0.8, 50, 0.95, a.frames, 1234, 0)
    eng.generate(prompt, trailing, pad, sp)                  # warm
    eng.timeline_arm(400000, a.cta)
    eng.generate(prompt, trailing, pad, sp)
    clk, tag = eng.timeline_read()
    print(f"entries {len(clk)}  total {(int(clk[-1]) - int(clk[0])) / a.mhz / 1e3:.2f} ms for prefill + {a.frames} frames "
          f"(generate_ms {eng.stats().last_generate_ms:.2f})")
    POINTS = {0: "begin", 1: "grid go", 2: "fetched", 3: "inputs staged", 4: "glue/dsmem push", 5: "softmax done", 6: "published", 7: "landed",
              8: "staged+csync", 9: "stages ready", 10: "mma done", 11: "stages released", 12: "partials csync", 13: "qkv normed", 14: "scores csync",
              15: "softmax csync"}
    # inside the sampler (phase kind 7) the same ids mark other places
    SPOINTS = {0: "begin", 1: "grid go", 7: "landed", 8: "logits loaded", 9: "max known", 10: "top-k bin found", 11: "candidates compacted",
               12: "exact top-k (warp 0)", 13: "softmax sum", 14: "top-p done", 15: "draw done", 3: "token broadcast", 4: "glue"}
    seg = collections.defaultdict(list)
    for i in range(1, len(clk)):
        st, kind, pt = (tag[i] >> 9) & 1, (tag[i] >> 4) & 31, tag[i] & 15
        seg[(st, kind, pt)].append(int(clk[i]) - int(clk[i - 1]))       # time spent reaching this point
    tot = sum(sum(v) for v in seg.values())
    print("each row: cycles from the previous mark to this point")
    print(f"{'stack':6s} {'phase':10s} {'->point':15s} {'count':>7s} {'mean cyc':>9s} {'mean us':>8s} {'share':>6s}")
    for k in sorted(seg):
        v = np.asarray(seg[k], np.float64)
        print(f"{'cp' if k[0] else 'talker':6s} {KINDS.get(k[1], str(k[1])):10s} {(SPOINTS if k[1] == 7 else POINTS).get(k[2], str(k[2])):20s} {len(v):7d} {v.mean():9.0f} "
              f"{v.mean() / a.mhz:8.2f} {v.sum() / tot:6.3f}")
    # the second recorder (lane 0 of another consumer warp): its lag behind thread 0 at the marks both pass, matched by occurrence
    clk2, tag2 = eng.timeline_read(1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "fk_timeline_raw.npz"), clk=clk, tag=tag, clk2=clk2, tag2=tag2)
    if len(clk2):
        occ = collections.defaultdict(list)
        for i in range(len(clk)):
            occ[int(tag[i])].append(int(clk[i]))
        occ2 = collections.defaultdict(list)
        for i in range(len(clk2)):
            occ2[int(tag2[i])].append(int(clk2[i]))
        print(f"second warp: {len(clk2)} entries; mean (t_warp2 - t_warp0) in cycles at common marks")
        for t in sorted(occ2):
            if t in occ and len(occ[t]) == len(occ2[t]):
                dlt = np.asarray(occ2[t], np.float64) - np.asarray(occ[t], np.float64)
                st, kind, pt = (t >> 9) & 1, (t >> 4) & 31, t & 15
                print(f"  {'cp' if st else 'talker':6s} {KINDS.get(kind, str(kind)):10s} {POINTS.get(pt, str(pt)):15s} n={len(dlt):5d} mean {dlt.mean():8.0f}  p10 {np.percentile(dlt, 10):8.0f}  p90 {np.percentile(dlt, 90):8.0f}")
    byp = collections.defaultdict(float)
    for k, v in seg.items():
        byp[POINTS.get(k[2], str(k[2]))] += sum(v)
    print("share by point:", {k: round(v / tot, 3) for k, v in sorted(byp.items(), key=lambda x: -x[1])})
    eng.close()


if __name__ == "__main__":
    main()
