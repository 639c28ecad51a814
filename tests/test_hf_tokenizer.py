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
    from tokenizers import Regex, Tokenizer, models, pre_tokenizers
    pattern = json.load(open(os.path.join(GOLD, "cases.json"), encoding="utf-8"))["pattern"]
    tok = Tokenizer(models.BPE.from_file(os.path.join(GOLD, "vocab.json"), os.path.join(GOLD, "merges.txt")))
    tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(Regex(pattern), behavior="isolated", invert=False),
                                                 pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
    rng = random.Random(11)
    alphabet = ("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789      \n\r\t'.,!?-_()[]{}@#$%&*+=/\\\"`~<>|^;:"
                "你好世界语音合成测试天气公园こんにちはテストカタカナ안녕하세요음성합성привестмирünïéàçñßÆø😀😃©®±×→∑∞𝒳١٢٣مرحباनमस्ते१२३\u00a0\u3000\u2028\u0085")
    texts = []
    for _ in range(3000):
        n = rng.randrange(1, 40)
        texts.append("".join(rng.choice(alphabet) for _ in range(n)))
    texts += ["'S", "'RE'LL", "a'd", "x's y'T", "'", "''", "'ſ"]
    got = run_hf(dump, texts, tmp_path)
    bad = [(t, tok.encode(t).ids, g) for t, g in zip(texts, got) if tok.encode(t).ids != g]
    assert not bad, bad[:3]
