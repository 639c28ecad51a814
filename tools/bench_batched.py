#!/usr/bin/env python
"""Exploration aid: frame-loop time of the batched path (lqt_synthesize_batch) for several slot counts / planes.
    python tools/bench_batched.py [--spec 0.6b] [--frames 64] [--batches 1,16,64,256] [--planes 3,2,1] [--vocode]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

load_package()
from leaxer_qwen3_tts_b200 import engine, modelspec as ms  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--spec", default="0.6b")
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--batches", default="1,16,64,256")
ap.add_argument("--planes", default="3,2,1")
ap.add_argument("--vocode", action="store_true")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
spec = {"0.6b": ms.spec_0p6b, "1.7b": ms.spec_1p7b, "tiny": ms.spec_tiny}[a.spec](0)
mdir = ms.generate_model_dir(ms.default_model_dir(spec), spec)
eng = engine.Engine(mdir, device=0, frame_impl="auto")
ids = engine.wrap_text_ids(ms.synthetic_text_ids(90, 1234))
for planes in [int(x) for x in a.planes.split(",")]:
    for B in [int(x) for x in a.batches.split(",")]:
        reqs = [{"token_ids": ids, "lang": "en", "utterance_id": u, "max_new_tokens": a.frames} for u in range(B)]
        best = None
        for rep in range(a.reps + 1):
            eng.reset_stats()
            t0 = time.perf_counter()
            outs = eng.synthesize_batch(reqs, 0.8, 50, 0.95, seed=1234, max_concurrent=B, planes=planes, vocode=a.vocode)
            wall = time.perf_counter() - t0
            st = eng.stats()
            if rep and (best is None or wall < best[0]):
                best = (wall, st.last_total_ms, st.last_frames, st.kernel_launches)
        wall, dev_ms, nfr, kl = best
        audio_s = sum(o[1].shape[0] for o in outs) * 0.08
        print(f"planes={planes} B={B:4d}: wall {wall * 1e3:8.1f} ms  device {dev_ms:8.1f} ms  {nfr} graph frames  "
              f"{dev_ms / max(nfr, 1):6.3f} ms/frame  {audio_s / wall:9.1f} audio-s/s  ({kl / max(nfr, 1):.0f} kernels/frame)", flush=True)
eng.close()
