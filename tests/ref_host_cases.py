"""Cases shared by tests/test_ref_host_pin.py and tests/golden/make_ref_host_golden.py: the reference's own host
(src/tts_onnx.cpp compiled unmodified against the ORT shim -> oracle/_ref/tts_host_ref) and the Python restatement
(oracle/qwen3_tts_oracle.py) are both run over the deterministic stub graphs (oracle/ort_shim/stub_graphs.h ==
oracle/stub_graphs.py) and must produce the same per-call trace."""
import hashlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import io_cases  # noqa: E402

HOST_REF = os.path.join(ROOT, "oracle", "_ref", "tts_host_ref")
IO_REF = os.path.join(ROOT, "oracle", "_ref", "io_dump_ref")
GRAPHS = ["text_project", "codec_embed", "code_predictor_embed", "talker_prefill", "talker_decode", "code_predictor",
          "tokenizer12hz_decode"]

IM_START, ASSISTANT, TTS_BOS, TTS_EOS, IM_END = 151644, 77091, 151672, 151673, 151645


def wrap(text_ids):
    return [IM_START, ASSISTANT, TTS_BOS] + [int(t) for t in text_ids] + [TTS_EOS, IM_END]


# name -> (lang, top_k, max_new, eos_at, text ids). top_k = 1 makes the reference's unseeded std::mt19937 draw deterministic.
ID_CASES = {
    "en_3text": ("en", 1, 4, -1, [1000, 2000, 3000]),
    "auto_1text": ("auto", 1, 3, -1, [77]),                      # P = 8, trailing = tts_eos only
    "zh_12text": ("zh", 1, 6, -1, [5, 151642, 9, 14990, 3, 8, 100000, 42, 7, 6, 5, 4]),   # trailing longer than the run
    "ko_eos": ("ko", 1, 9, 9 + 3, [11, 12]),                     # the decode stub favours CODEC_EOS at mask length 12: 3 frames
    "ja_pad_after_trailing": ("ja", 1, 5, -1, [21, 22]),         # frames beyond trailing_len add tts_pad (:833-842)
}
TEXT_CASES = {"text_hello_world": ("en", 1, 3, -1, "hello world")}       # through the reference's tokenizer (synthetic vocab)
CLONE_CASES = {"clone_zh": ("zh", 1, 3, -1, "speech testing 123")}       # P = 10: speaker row before codec_bos (:481-490)


def make_model_dir(base, with_speaker=True, with_tokenizer=True):
    """An empty file per graph is all the reference's loader looks at (fs::exists, src/tts_onnx.cpp:136); the tokenizer
    files go where the constructor expects them (:110-112)."""
    mdir = os.path.join(base, "onnx_kv")
    os.makedirs(mdir, exist_ok=True)
    for g in GRAPHS + (["speaker_encoder"] if with_speaker else []):
        open(os.path.join(mdir, g + ".onnx"), "wb").close()
    if with_tokenizer:
        io_cases.write_tokenizer_files(os.path.join(base, "models", "Qwen3-TTS-12Hz-0.6B-Base"))
    return mdir


def write_ref_wav(path, seconds=3.0, rate=24000):
    """BASELINE config 3: synthetic 3 s 24 kHz 16-bit mono clip (three sines + seeded noise)."""
    n = int(seconds * rate)
    t = np.arange(n) / rate
    x = 0.4 * np.sin(2 * np.pi * 220 * t) + 0.25 * np.sin(2 * np.pi * 1330 * t) + 0.1 * np.sin(2 * np.pi * 5100 * t)
    x = x + 0.02 * np.random.default_rng(3).standard_normal(n)
    io_cases._wav(path, 1, 1, rate, 16, (np.clip(x, -0.99, 0.99) * 32767).astype("<i2").tobytes())
    return path


def run_ref(args):
    """-> (trace lines, result line)"""
    out = subprocess.run([HOST_REF, *[str(a) for a in args]], check=True, stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL).stdout.decode()
    lines = out.strip().split("\n")
    assert lines[0].startswith("READY"), lines[:2]
    assert lines[-1].startswith("RESULT"), lines[-1]
    return lines[1:-1], lines[-1]


def sha(lines):
    return hashlib.sha256("\n".join(lines).encode()).hexdigest()


def result_line(audio: np.ndarray, calls: int) -> str:
    from oracle import stub_graphs as sg
    s, x = sg.digest_words(np.ascontiguousarray(audio, dtype=np.float32))
    return f"RESULT audio {audio.shape[0]} {s:08x} {x:08x} calls {calls}"


def ref_melwav(binary, wav):
    b = subprocess.run([binary, "melwav", wav], check=True, stdout=subprocess.PIPE).stdout
    frames = int(np.frombuffer(b[:4], "<i4")[0])
    return np.frombuffer(b[4:], "<f4").reshape(128, frames).copy()


FILTER_CASES = [(3072, 50, 0.95, 0), (2048, 50, 0.95, 1), (2048, 5, 0.5, 2), (3072, 1, 0.95, 3), (2048, 200, 0.9, 4),
                (2048, 50, 0.3, 5), (3072, 2047, 0.999, 6)]


def filter_logits(V, seed, ties=False):
    x = (np.random.default_rng(100 + seed).standard_normal(V) * 3.2).astype(np.float32)
    if ties:
        x = (np.round(x * 2) / 2).astype(np.float32)
    return x
