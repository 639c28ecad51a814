#!/usr/bin/env python
"""Golden vectors at the BENCHMARKED shape (BASELINE.json configs[1], C2): full 0.6B random-init model (seed 0),
90 synthetic text ids, en, 375 frames (30 s), temp 0.8 / top-k 50 / top-p 0.95, Philox (1234, utterance 0), written by the
CPU oracle (schedule "cached": arithmetic-identical to the reference's cache-less schedule, tests/test_cpu_oracle.py) so that
the GPU box does not pay ~10 minutes of CPU time per run:
  codes            [375,16] free-running, bf16 talker KV (the engine's default / the benchmarked mode)
  codes_f32kv      [375,16] free-running, fp32 talker KV (the engine's parity mode, LQT_KV_F32)
  frames           indices of the frames whose logits are stored
  talker_logits    [len(frames), 3072]   (masked specials = -inf), from the bf16-KV run
  cp_logits        [len(frames), 15, 2048]
Used by tests/test_gpu_parity.py::test_c2_* and by bench.py's parity check. Run: python tests/golden/make_c2_golden.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import qwen3_tts_oracle as orc  # noqa: E402
from leaxer_qwen3_tts_b200 import modelspec as ms  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FRAMES = int(os.environ.get("C2_FRAMES", "375"))
SEL = [f for f in (0, 1, 2, 31, 54, 55, 63, 64, 127, 128, 191, 255, 319, 373, 374) if f < FRAMES]


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    spec = ms.spec_0p6b(0)
    mdir = ms.generate_model_dir(ms.default_model_dir(spec), spec)
    ids = orc.wrap_text_ids(orc.synthetic_text_ids(90, 1234))
    sp = orc.SamplingParams(temperature=0.8, top_k=50, top_p=0.95, max_new_tokens=FRAMES, seed=1234, utterance_id=0)
    out = {"token_ids": np.asarray(ids, np.int64), "frames": np.asarray(SEL, np.int32)}
    for tag, kv_bf16 in (("", True), ("_f32kv", False)):
        m = orc.OracleModel(mdir, kv_bf16=kv_bf16)
        tr = {}
        t0 = time.time()
        _, codes = orc.synthesize_tokens(m, ids, "en", sp, trace=tr, run_vocoder=False)
        print(f"kv_bf16={kv_bf16}: {codes.shape[0]} frames in {time.time() - t0:.0f}s", flush=True)
        out["codes" + tag] = codes
        if kv_bf16:
            out["talker_logits"] = np.stack([tr["talker_logits"][f] for f in SEL]).astype(np.float32)
            out["cp_logits"] = np.stack([tr["cp_logits"][f] for f in SEL]).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "c2_full_375.npz"), **out)
    print("wrote c2_full_375.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
