"""Shared case definitions for the io/ parity tests (tests/test_host_io.py) and the fixture generator
(tests/golden/make_io_golden.py): synthetic WAV files covering the reference's reader branches
(src/io/wav_reader.cpp:29-140; cases modelled on the reference's tests/test_wav_reader.cpp and tests/test_mel.cpp),
a synthetic BPE vocabulary (the real Qwen vocab.json / merges.txt are not available offline) and the command lines
for the io_dump driver (oracle/io_dump.cpp)."""
import json
import os
import struct

import numpy as np


def _wav(path, tag, channels, rate, bits, payload, fmt_extra=b"", pre_chunks=b"", riff=b"RIFF", wave=b"WAVE"):
    fmt = struct.pack("<HHIIHH", tag, channels, rate, rate * channels * bits // 8, channels * bits // 8, bits) + fmt_extra
    body = wave + pre_chunks + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(payload)) + payload
    with open(path, "wb") as f:
        f.write(riff + struct.pack("<I", len(body)) + body)


def write_wav_cases(d):
    """-> list of (name, path)"""
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(5)
    x = np.clip(rng.standard_normal(1000) * 0.3, -0.99, 0.99)
    cases = []

    def add(name, *a, **k):
        p = os.path.join(d, name + ".wav")
        _wav(p, *a, **k)
        cases.append((name, p))

    add("s16_mono_24k", 1, 1, 24000, 16, (x * 32767).astype("<i2").tobytes())
    add("s16_stereo_44k", 1, 2, 44100, 16, (np.stack([x, -0.5 * x], 1) * 32767).astype("<i2").tobytes())
    add("u8_mono", 1, 1, 8000, 8, ((x * 127) + 128).astype(np.uint8).tobytes())
    s24 = (x * 8388607).astype(np.int32)
    add("s24_mono", 1, 1, 48000, 24, b"".join(struct.pack("<i", int(v))[:3] for v in s24))
    add("s32_mono", 1, 1, 16000, 32, (x * 2147483647).astype("<i4").tobytes())
    add("f32_mono", 3, 1, 24000, 32, x.astype("<f4").tobytes())
    add("f64_mono_silence", 3, 1, 24000, 64, x.astype("<f8").tobytes())
    add("fmt18_extra", 1, 1, 22050, 16, (x * 32767).astype("<i2").tobytes(), fmt_extra=b"\x00\x00")
    add("list_chunk_first", 1, 1, 24000, 16, (x * 32767).astype("<i2").tobytes(),
        pre_chunks=b"LIST" + struct.pack("<I", 10) + b"INFOabcdef")
    add("extensible_rejected", 0xFFFE, 1, 24000, 16, (x * 32767).astype("<i2").tobytes(), fmt_extra=b"\x16\x00" + b"\x00" * 22)
    add("bad_riff", 1, 1, 24000, 16, b"\x00\x00", riff=b"RIFX")
    add("zero_channels", 1, 0, 24000, 16, b"\x00\x00")
    p = os.path.join(d, "truncated_data.wav")                       # data chunk longer than the file: missing samples stay zero
    _wav(p, 1, 1, 24000, 16, (x * 32767).astype("<i2").tobytes())
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-600])
    cases.append(("truncated_data", p))
    cases.append(("missing_file", os.path.join(d, "does_not_exist.wav")))
    return cases


MEL_CASES = [(72000, 1), (24000, 2), (1024, 3), (700, 4), (1279, 5), (1280, 6)]          # 3 s = 278 frames; < win = 1 frame
RESAMPLE_CASES = [(1000, 16000, 24000), (1000, 44100, 24000), (777, 24000, 24000), (500, 48000, 24000), (3, 8000, 24000)]


def write_tokenizer_files(d):
    """A small byte-level BPE vocabulary in the real files' format: vocab.json (flat object, with \\u escapes) and
    merges.txt (with the '#version' header line the reference stores as a harmless rank-0 pair)."""
    os.makedirs(d, exist_ok=True)
    toks = ["h", "e", "l", "o", "w", "r", "d", "s", "p", "c", "t", "i", "n", "y", "g", "a", "Ġ", "!", ",", "1", "2", "3",
            "he", "ll", "hell", "hello", "Ġw", "or", "Ġwor", "ld", "Ġworld", "sp", "ee", "ch", "Ġsp", "in", "ing", "12",
            "ł", "ĠĠ", "'", "'s", "Ċ"]
    vocab = {t: 1000 + i for i, t in enumerate(toks)}
    merges = ["#version: 0.2", "h e", "l l", "he ll", "hell o", "Ġ w", "o r", "Ġw or", "l d", "Ġwor ld", "s p", "e e",
              "c h", "Ġ sp", "i n", "in g", "1 2", "Ġ Ġ", "' s"]
    vp, mp = os.path.join(d, "vocab.json"), os.path.join(d, "merges.txt")
    with open(vp, "w", encoding="utf-8") as f:
        f.write(json.dumps(vocab, ensure_ascii=True))               # non-ASCII tokens as \uXXXX escapes, like the real file
    with open(mp, "w", encoding="utf-8") as f:
        f.write("\n".join(merges) + "\n")
    return vp, mp


TOKENIZER_TEXTS = ["hello world", "hello  world!", "speech testing 123", "a_b it's", "hello\nworld", "你好 hello", "",
                   "Hello, WORLD", "   ", "12 3"]
