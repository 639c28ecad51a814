#!/usr/bin/env python
"""Runs the vocoder alone (lqt_vocoder_decode, T frames of random codes) a few times: for ncu launch lists / timing."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
load_package()
from leaxer_qwen3_tts_b200 import engine, modelspec as ms
T = int(sys.argv[1]) if len(sys.argv) > 1 else 375
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
spec = ms.spec_0p6b(0)
eng = engine.Engine(ms.generate_model_dir(ms.default_model_dir(spec), spec), device=0)
codes = np.random.default_rng(0).integers(0, 2048, size=(T, 16))
for i in range(reps):
    eng.vocoder_decode(codes)
    print(f"vocoder T={T}: {eng.stats().last_vocoder_ms:.2f} ms", flush=True)
eng.close()
