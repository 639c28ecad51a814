#!/usr/bin/env python
"""Generates tests/golden/io_reference.npz from the REFERENCE's io sources compiled by oracle/Makefile
(oracle/_ref/io_dump_ref <- /root/reference/src/io/{mel,wav_reader,tokenizer}.cpp). Run in the container that has
/root/reference mounted:  make -C oracle && python tests/golden/make_io_golden.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import io_cases  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "io_dump_ref")


def run(binary, *args):
    return subprocess.run([binary, *[str(a) for a in args]], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout


def collect(binary, workdir):
    out = {}
    for n, seed in io_cases.MEL_CASES:
        b = run(binary, "mel", n, seed)
        out[f"mel_{n}_{seed}_frames"] = np.frombuffer(b[:4], "<i4").copy()
        out[f"mel_{n}_{seed}"] = np.frombuffer(b[4:], "<f4").copy()
    for name, path in io_cases.write_wav_cases(os.path.join(workdir, "wav")):
        b = run(binary, "wav", path)
        out[f"wav_{name}_sr"] = np.frombuffer(b[:4], "<i4").copy()
        out[f"wav_{name}"] = np.frombuffer(b[4:], "<f4").copy()
    for n, s, d in io_cases.RESAMPLE_CASES:
        out[f"resample_{n}_{s}_{d}"] = np.frombuffer(run(binary, "resample", n, s, d), "<f4").copy()
    vp, mp = io_cases.write_tokenizer_files(os.path.join(workdir, "tok"))
    for i, text in enumerate(io_cases.TOKENIZER_TEXTS):
        out[f"tok_{i}"] = np.frombuffer(run(binary, "tok", vp, mp, text), "<i4").copy()
        out[f"tok_novocab_{i}"] = np.frombuffer(run(binary, "tok", "-", "-", text), "<i4").copy()
        out[f"tok_nomerges_{i}"] = np.frombuffer(run(binary, "tok", vp, "-", text), "<i4").copy()
    return out


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as d:
        data = collect(REF, d)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "io_reference.npz"), **data)
    print(f"wrote {len(data)} arrays from {REF}")
