#!/usr/bin/env python
"""One launch of the persistent frame kernel for an ncu capture (BASELINE configs[1] prompt, --frames frames, seeded sampling):
LQT_FK_NOCOOP=1 ncu --set full --import-source on --clock-control none -k regex:frame_kernel -c 1 -o X python tools/fk_ncu_run.py --frames 8"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

load_package()
from leaxer_qwen3_tts_b200 import engine, modelspec as ms  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=8)
a = ap.parse_args()
spec = ms.spec_0p6b(0)
mdir = ms.generate_model_dir(ms.default_model_dir(spec), spec)
eng = engine.Engine(mdir)
prompt, trailing, pad = eng.build_prompt(engine.wrap_text_ids(ms.synthetic_text_ids(90, 1234)), "en")
codes = eng.generate(prompt, trailing, pad, eng.sampling(0.8, 50, 0.95, a.frames, 1234, 0))
print("frames", len(codes), "generate_ms", eng.stats().last_generate_ms)
eng.close()
