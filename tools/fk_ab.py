#!/usr/bin/env python
"""A/B of frame-kernel builds on the GPU box: for every library given, one process runs BASELINE configs[1]'s generation
(0.6B, 375 frames, seeded top-k 50 / top-p 0.95) through lqt_generate, prints the frame-loop time per frame (best and median
of --reps) and checks the codes against the CPU oracle's golden (tests/golden/c2_full_375.npz): bf16 KV must agree for at
least the golden's known-exact prefix, fp32 KV (--f32) for all frames.
Usage: python tools/fk_ab.py [--frames 375] [--reps 5] [--f32] lib1.so [lib2.so ...]"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(a):
    sys.path.insert(0, ROOT)
    import numpy as np
    from __graft_entry__ import load_package
    load_package()
    from leaxer_qwen3_tts_b200 import engine, modelspec as ms
    spec = ms.spec_0p6b(0)
    mdir = ms.generate_model_dir(ms.default_model_dir(spec), spec)
    g = np.load(os.path.join(ROOT, "tests", "golden", "c2_full_375.npz"))
    out = {"lib": os.path.basename(os.environ.get("LQT_B200_LIB", "default"))}
    for kv in (["bf16", "f32"] if a.f32 else ["bf16"]):
        eng = engine.Engine(mdir, kv_dtype=kv)
        prompt, trailing, pad = eng.build_prompt(g["token_ids"], "en")
        sp = eng.sampling(0.8, 50, 0.95, a.frames, 1234, 0)
        codes = eng.generate(prompt, trailing, pad, sp)
        ts = []
        for _ in range(a.reps):
            c2 = eng.generate(prompt, trailing, pad, sp)
            assert (c2 == codes).all(), "run-to-run difference"
            ts.append(eng.stats().last_generate_ms)
        ref = g["codes_f32kv" if kv == "f32" else "codes"][: a.frames]
        n = min(len(codes), len(ref))
        neq = np.nonzero((codes[:n] != ref[:n]).any(axis=1))[0]
        out[kv] = {"ms_per_frame_best": min(ts) / max(1, len(codes)), "ms_per_frame_med": float(np.median(ts)) / max(1, len(codes)),
                   "frames": int(len(codes)), "exact_prefix": int(neq[0]) if len(neq) else int(n)}
        eng.close()
    print("FKAB " + json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=375)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--f32", action="store_true")
    ap.add_argument("--child", action="store_true")
    ap.add_argument("libs", nargs="*")
    a = ap.parse_args()
    if a.child:
        return child(a)
    for lib in a.libs or [os.path.join(ROOT, "leaxer-qwen3-tts_b200", "csrc", "liblqt_b200.so")]:
        env = dict(os.environ, LQT_B200_LIB=os.path.abspath(lib))
        cmd = [sys.executable, os.path.abspath(__file__), "--child", "--frames", str(a.frames), "--reps", str(a.reps)] + (["--f32"] if a.f32 else [])
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        lines = [l for l in r.stdout.splitlines() if l.startswith("FKAB ")]
        print(lines[-1] if lines else f"FKAB {{\"lib\": \"{os.path.basename(lib)}\", \"error\": {json.dumps((r.stderr or r.stdout)[-400:])}}}", flush=True)


if __name__ == "__main__":
    main()
