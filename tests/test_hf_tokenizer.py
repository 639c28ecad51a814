"""HF-faithful tokenizer mode of the host (leaxer-qwen3-tts_b200/host/io/tokenizer.cpp) against the HuggingFace `tokenizers`
library: committed golden ids (tests/golden/hf_tok, made by make_hf_tokenizer_golden.py with the published Qwen2 pipeline on a small
trained vocabulary) and, when the library is importable, a live comparison on random multilingual strings. The DEFAULT mode stays the
reference's (quirks included) -- that one is checked against the reference's own sources in tests/test_host_io.py."""
import json
import os
import random
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "hf_tok")
HOST = os.path.join(ROOT, "leaxer-qwen3-tts_b200", "host")


@pytest.fixture(scope="module")
def dump():
    r = subprocess.run(["make", "-C", HOST, "build/io_dump"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return os.path.join(HOST, "build", "io_dump")


def escape(t: str) -> str:
    return t.replace("\\", "\\\\").replace("\n", "\\n").replace("\r", "\\r").replace("\t", "\\t")


def run_hf(dump, texts, tmp_path, vocab=None, merges=None):
    f = tmp_path / "texts.txt"
    f.write_bytes("".join(escape(t) + "\n" for t in texts).encode("utf-8"))
    out = subprocess.run([dump, "tokhf", vocab or os.path.join(GOLD, "vocab.json"), merges or os.path.join(GOLD, "merges.txt"), str(f)],
                         check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    a = np.frombuffer(out, "<i4")
    res, p = [], 0
    for _ in texts:
        n = int(a[p]); res.append(a[p + 1:p + 1 + n].tolist()); p += 1 + n
    assert p == a.size
    return res


def test_golden_ids(dump, tmp_path):
    cases = json.load(open(os.path.join(GOLD, "cases.json"), encoding="utf-8"))["cases"]
    got = run_hf(dump, [c["text"] for c in cases], tmp_path)
    bad = [(c["text"], c["ids"], g) for c, g in zip(cases, got) if c["ids"] != g]
    assert not bad, bad[:3]


def test_reference_mode_is_still_the_default(dump, tmp_path):
    """the default keeps the reference's behaviour: the bytes of a CJK character are never merged (bytes >= 161 stay raw single
    bytes that match nothing), one id per byte, most of them the raw byte value; HF mode merges them into vocabulary tokens"""
    out = subprocess.run([dump, "tok", os.path.join(GOLD, "vocab.json"), os.path.join(GOLD, "merges.txt"), "你好"], check=True,
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env={k: v for k, v in os.environ.items() if k != "LEAXER_TOKENIZER"}).stdout
    ids = np.frombuffer(out, "<i4")[1:].tolist()
    raw = list("你好".encode("utf-8"))
    assert len(ids) == len(raw) and sum(a == b for a, b in zip(ids, raw)) >= 5
    hf = run_hf(dump, ["你好"], tmp_path)[0]
    assert len(hf) < len(raw) and hf != ids


def test_live_against_the_tokenizers_library(dump, tmp_path):
    tk = pytest.importorskip("tokenizers")
    from tokenizers import Regex, Tokenizer, models, normalizers, pre_tokenizers
    pattern = json.load(open(os.path.join(GOLD, "cases.json"), encoding="utf-8"))["pattern"]
    tok = Tokenizer(models.BPE.from_file(os.path.join(GOLD, "vocab.json"), os.path.join(GOLD, "merges.txt")))
    tok.normalizer = normalizers.NFC()
    tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(Regex(pattern), behavior="isolated", invert=False),
                                                 pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
    rng = random.Random(11)
    alphabet = ("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789      \n\r\t'.,!?-_()[]{}@#$%&*+=/\\\"`~<>|^;:"
                "你好世界语音合成测试天气公园こんにちはテストカタカナ안녕하세요음성합성привестмирünïéàçñßÆø😀😃©®±×→∑∞𝒳١٢٣مرحباनमस्ते१२३\u00a0\u3000\u2028\u0085"
                "\u0301\u0308\u0323\u0307\u030a\u0327\u1100\u1112\u1161\u1173\u11ab\u11af\u3099\u309a\u093c\u212b\u2126\ufb01")   # not NFC on input
    texts = []
    for _ in range(3000):
        n = rng.randrange(1, 40)
        texts.append("".join(rng.choice(alphabet) for _ in range(n)))
    texts += ["'S", "'RE'LL", "a'd", "x's y'T", "'", "''", "'ſ"]
    got = run_hf(dump, texts, tmp_path)
    bad = [(t, tok.encode(t).ids, g) for t, g in zip(texts, got) if tok.encode(t).ids != g]
    assert not bad, bad[:3]


def run_nfc(dump, texts, tmp_path):
    """str items: one escaped line each; bytes items: written as they are (raw lines, '\\n' included)"""
    f = tmp_path / "nfc.txt"
    f.write_bytes(b"".join(t if isinstance(t, bytes) else escape(t).encode("utf-8") + b"\n" for t in texts))
    out = subprocess.run([dump, "nfc", str(f)], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    return out.split(b"\n")[:-1]


def test_nfc_matches_unicodedata(dump, tmp_path):
    """normalize_nfc (the first stage of HF mode) against Python's unicodedata -- the tables were generated from the same Unicode
    version: every code point with a canonical decomposition or a non-zero combining class alone and behind a base letter, conjoining
    jamo, random mixtures with marks out of canonical order."""
    import unicodedata
    rng = random.Random(3)
    pool = [chr(cp) for cp in range(0x110000) if not 0xD800 <= cp <= 0xDFFF and
            ((unicodedata.decomposition(chr(cp)) and not unicodedata.decomposition(chr(cp)).startswith("<")) or unicodedata.combining(chr(cp)))]
    pool += [chr(c) for c in range(0x1100, 0x1113)] + [chr(c) for c in range(0x1161, 0x1176)] + [chr(c) for c in range(0x11A7, 0x11C3)]
    pool += [chr(rng.randrange(0xAC00, 0xD7A4)) for _ in range(100)]
    base = "aeouAEINcsz '.1你こカ한ㄱ"
    texts = pool + ["a" + c for c in pool]
    for _ in range(4000):
        texts.append("".join(rng.choice(pool) if rng.random() < 0.6 else rng.choice(base) for _ in range(rng.randrange(1, 12))))
    texts += ["A\u030a", "\u212b", "\u1e0b\u0323", "D\u0323\u0307", "\u1100\u1161\u11a8", "\uac00\u11a8", "\u0958", "\u0f73\u0f71", "\u09c7\u09be"]
    got = run_nfc(dump, texts, tmp_path)
    assert len(got) == len(texts)
    bad = [(t, g) for t, g in zip(texts, got) if unicodedata.normalize("NFC", t).encode("utf-8") != g]
    assert not bad, bad[:3]


def test_nfc_leaves_invalid_bytes_alone(dump, tmp_path):
    """bytes that are not UTF-8 pass through unchanged and separate the runs that are normalised (the reference's tokenizer never
    rejects input either); ASCII / Latin-1 text is returned as it is"""
    raw = [b"plain ascii\n", b"caf\xc3\xa9 latin-1 range\n", b"e\xcc\x81 \xff e\xcc\x81\xfe\xcc\x81 \xed\xa0\x80 \xcc\n", b"\xe1\x84\x80\xe1\x85\xa1\xc0\xe1\x86\xa8\n"]
    got = run_nfc(dump, raw, tmp_path)
    assert got[0] == b"plain ascii" and got[1] == b"caf\xc3\xa9 latin-1 range"
    assert got[2] == b"\xc3\xa9 \xff \xc3\xa9\xfe\xcc\x81 \xed\xa0\x80 \xcc"          # the lone mark behind the invalid byte stays a mark
    assert got[3] == b"\xea\xb0\x80\xc0\xe1\x86\xa8"                                  # L + V compose; the trailing T is cut off by the invalid byte
