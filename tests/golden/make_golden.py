"""Generates the committed golden fixtures from the CPU oracle on the full 0.6B random-init model
(seed 0). Run here (CPU container): python tests/golden/make_golden.py
The reference itself cannot produce vectors (no ONNX Runtime, no .onnx graphs: SURVEY.md §8c), so
these pin the oracle <-> CUDA-engine contract, not the reference's numbers ("parity unpinned")."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import qwen3_tts_oracle as orc  # noqa: E402
from leaxer_qwen3_tts_b200 import modelspec as ms  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = ms.spec_0p6b(0)
    mdir = ms.generate_model_dir(ms.default_model_dir(spec), spec, verbose=True)
    m = orc.OracleModel(mdir)
    # C1: "Hello world", en, greedy, 25 frames (2 s). Two fixed text ids stand in for the tokenizer
    # output (vocab.json is not available offline).
    ids = orc.wrap_text_ids([9707, 1879])
    tr = {}
    audio, codes = orc.synthesize_tokens(m, ids, "en", orc.SamplingParams(max_new_tokens=25, greedy=True), trace=tr)
    np.savez_compressed(
        os.path.join(HERE, "c1_hello_world_en_greedy.npz"),
        token_ids=np.asarray(ids, np.int64), codes=codes, audio=audio.astype(np.float32),
        prompt=tr["prompt"], trailing=tr["trailing"], tts_pad=tr["tts_pad"],
        talker_logits_f0=tr["talker_logits"][0], talker_logits_f24=tr["talker_logits"][24],
        last_hidden_f0=tr["last_hidden"][0], cp_logits_f0=tr["cp_logits"][0], next_in_f0=tr["next_in"][0])
    # C2-shaped, shortened: 90 synthetic text ids, seeded sampler, 8 frames
    ids2 = orc.wrap_text_ids(orc.synthetic_text_ids(90, 1234))
    sp = orc.SamplingParams(temperature=0.8, top_k=50, top_p=0.95, max_new_tokens=8, seed=1234, utterance_id=0)
    _, codes2 = orc.synthesize_tokens(m, ids2, "en", sp, run_vocoder=False)
    np.savez_compressed(os.path.join(HERE, "c2_short_seeded.npz"), token_ids=np.asarray(ids2, np.int64), codes=codes2)
    # tiny-spec vocoder + sampler vectors (cheap to re-check on CPU)
    ts = ms.spec_tiny(0)
    tm = orc.OracleModel(ms.generate_model_dir(ms.default_model_dir(ts), ts))
    tcodes = np.random.default_rng(0).integers(0, 2048, size=(5, 16))
    ta, _ = tm.vocoder(tcodes)
    lg = (np.random.default_rng(1).standard_normal((8, 3072)) * 3.2).astype(np.float32)
    toks = [orc.sample_token(lg[i], orc.SamplingParams(seed=1234, utterance_id=i), i, i) for i in range(8)]
    np.savez_compressed(os.path.join(HERE, "tiny_vectors.npz"), codes=tcodes, audio=ta.numpy(),
                        sampler_logits=lg, sampler_tokens=np.asarray(toks, np.int64))
    print("golden written:", os.listdir(HERE))


if __name__ == "__main__":
    main()
