#!/usr/bin/env python
"""Golden vectors for the HF-faithful tokenizer mode (host/io/tokenizer.cpp), made with the HuggingFace `tokenizers`
library -- the implementation the published Qwen2/Qwen3 tokenizer.json runs on. No real vocabulary exists offline, so a small
byte-level BPE is TRAINED here on a multilingual corpus with exactly the published pipeline (NFC normaliser; Split on the Qwen2
pattern, isolated; ByteLevel without prefix space and without its own regex); what is pinned is the pipeline -- normalisation,
pre-tokenisation, byte alphabet, merge order -- not the vocabulary. Writes tests/golden/hf_tok/{vocab.json, merges.txt, cases.json}."""
import json
import os
import random

from tokenizers import Regex, Tokenizer, models, normalizers, pre_tokenizers, trainers

QWEN2_PATTERN = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+")

CORPUS = [
    "Hello world, this is a speech synthesis test. It's working, isn't it? We've tested it; they'll say they'd like more.",
    "The quick brown fox jumps over the lazy dog 1234567890 times, at 3.14159 km/h (approx.)!",
    "你好,世界!这是一个语音合成测试。今天天气很好,我们一起去公园散步吧。",
    "こんにちは世界。これは音声合成のテストです。今日はいい天気ですね。カタカナとひらがな。",
    "안녕하세요 세계. 이것은 음성 합성 테스트입니다. 오늘 날씨가 좋네요.",
    "Привет, мир! Это тест синтеза речи. Сегодня хорошая погода.",
    "Ünïcödé tëxt wíth àccénts: naïve café, façade, jalapeño, Ærøskøbing, Straße.",
    "Emoji 😀😃 and symbols ©®™ ± × ÷ → ← ∑ ∫ √ ∞ ≠ ≤ ≥ and math 𝒳𝒴𝒵.",
    "Tabs\tand\nnewlines\r\n\r\n   multiple   spaces   \n\n  trailing  ",
    "email@example.com http://example.org/path?q=1&r=2 #hashtag @mention $100 50% C++ a_b_c",
    "مرحبا بالعالم. هذا اختبار تركيب الكلام. ١٢٣٤٥",
    "नमस्ते दुनिया। यह एक भाषण संश्लेषण परीक्षण है। १२३",
]

CASES = [
    "hello", "world", "Hello world", " leading space", "trailing space ", "  two  spaces  ", "it's", "IT'S", "we'Re", "they'll've", "'sample",
    "1234567890", "3.14159", "a1b2c3", "你好世界", "你好,世界!", "こんにちは世界。", "안녕하세요 세계", "Привет, мир!", "naïve café", "Straße",
    "😀😃", "𝒳𝒴𝒵 math", "©®™", "tab\there", "line\nbreak", "crlf\r\nhere", "\n\n\n", "  \n  x", "x \n", "x  ", "x   y", " ", "  ", "\t",
    "a_b_c", "C++", "$100", "50%", "http://example.org/path?q=1", "mixed 中文 and English 123", "!!!\n", "?!\r\n\r\nnext", " !x", "  !x",
    "٣ مرحبا", "नमस्ते १२३", " nbsp　ideographic", " ls", "á combining", "",
    # not in NFC on input: combining sequences, conjoining jamo, singletons (the Angstrom / Kelvin / Ohm signs), marks out of canonical
    # order, a composition exclusion (stays decomposed), a compatibility ligature (NFC leaves it alone)
    "cafe\u0301 nai\u0308ve", "\u1112\u1161\u11ab\u1100\u1173\u11af", "\u212b \u212a \u2126", "q\u0307\u0323 d\u0323\u0307", "\u0915\u093c \u0958",
    "\ufb01n", "A\u030a\u0301", "\u30ab\u3099\u30cf\u309a", "o\u0302\u0303 \u1ed7", "\u0627\u0653",
]


def main():
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hf_tok")
    os.makedirs(out, exist_ok=True)
    tok = Tokenizer(models.BPE())
    tok.normalizer = normalizers.NFC()                    # tokenizer.json of the published checkpoints: "normalizer": {"type": "NFC"}
    tok.pre_tokenizer = pre_tokenizers.Sequence([
        pre_tokenizers.Split(Regex(QWEN2_PATTERN), behavior="isolated", invert=False),
        pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
    trainer = trainers.BpeTrainer(vocab_size=900, initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), special_tokens=[], show_progress=False)
    tok.train_from_iterator(CORPUS * 3, trainer)
    tok.model.save(out)
    rng = random.Random(5)
    pool = "".join(CORPUS)
    cases = list(CASES)
    for _ in range(40):                                   # random substrings of the corpus (code-point boundaries)
        a = rng.randrange(len(pool)); b = min(len(pool), a + rng.randrange(1, 40))
        cases.append(pool[a:b])
    rec = [{"text": t, "ids": tok.encode(t).ids} for t in cases]
    with open(os.path.join(out, "cases.json"), "w", encoding="utf-8") as f:
        json.dump({"pattern": QWEN2_PATTERN, "cases": rec}, f, ensure_ascii=True, indent=0)
    print(f"{out}: vocab {tok.get_vocab_size()}, {len(rec)} cases")


if __name__ == "__main__":
    main()
