""".lqw weight files: the loader's header checks (csrc/lqw_loader.h parse_lqw_header) through the C-ABI entry point
lqt_check_model_file, which needs no GPU. A corrupt or hostile file must be rejected with a reason before any device
allocation -- never read past the header block, never accept a tensor whose byte count disagrees with its dims or that lies
outside the file (lqw_loader.h:77). The dtype/shape check of every tensor against the model spec happens in
lqt_create (engine.cu need<T>) and is covered on the GPU by tests/test_gpu_parity.py::test_create_rejects_wrong_shapes."""
import os
import struct

import numpy as np
import pytest

from leaxer_qwen3_tts_b200 import engine, modelspec as ms


@pytest.fixture(scope="module")
def good(tmp_path_factory):
    d = tmp_path_factory.mktemp("lqw")
    p = str(d / "t.lqw")
    ms.write_lqw(p, [("a", ms.DT_F32, np.arange(12, dtype=np.float32).reshape(3, 4)),
                     ("w", ms.DT_BF16, ms.f32_to_bf16_bits(np.ones((8, 16), np.float32)))], {"hidden": "16", "graph": "x"})
    return p


def _mutate(src, dst, fn):
    raw = bytearray(open(src, "rb").read())
    fn(raw)
    open(dst, "wb").write(raw)
    return dst


def test_well_formed_file_passes(good, tiny_dir):
    assert engine.check_model_file(good) == ""
    for g in ms.GRAPH_FILES:
        assert engine.check_model_file(os.path.join(tiny_dir, g + ".lqw")) == ""
    meta, tensors = ms.read_lqw(good)
    assert meta["hidden"] == "16" and tensors["a"].shape == (3, 4)


def test_missing_and_bad_magic(good, tmp_path):
    assert "cannot open" in engine.check_model_file(str(tmp_path / "nope.lqw"))
    bad = _mutate(good, str(tmp_path / "m.lqw"), lambda r: r.__setitem__(slice(0, 4), b"XXXX"))
    assert "bad magic" in engine.check_model_file(bad)
    open(str(tmp_path / "short.lqw"), "wb").write(b"LQTW0001\x01")
    assert "bad magic" in engine.check_model_file(str(tmp_path / "short.lqw"))


def test_counts_and_strings_cannot_run_past_the_header(good, tmp_path):
    # tensor count far larger than the table that follows
    p = _mutate(good, str(tmp_path / "nt.lqw"), lambda r: r.__setitem__(slice(8, 12), struct.pack("<I", 5000)))
    assert "corrupt header" in engine.check_model_file(p)       # (the zero padding behind the table parses as a bad entry first)
    p = _mutate(good, str(tmp_path / "nt2.lqw"), lambda r: r.__setitem__(slice(8, 12), struct.pack("<I", 0xFFFFFFF0)))
    assert "bad header" in engine.check_model_file(p)
    # first meta key length = 0xFFFF
    p = _mutate(good, str(tmp_path / "kl.lqw"), lambda r: r.__setitem__(slice(24, 26), struct.pack("<H", 0xFFFF)))
    assert "runs past the header" in engine.check_model_file(p)
    # data_start beyond the file / absurd
    p = _mutate(good, str(tmp_path / "ds.lqw"), lambda r: r.__setitem__(slice(16, 24), struct.pack("<Q", 1 << 40)))
    assert "bad header" in engine.check_model_file(p)


def _tensor_table_pos(raw):
    """offset of the first tensor entry (after the meta pairs)"""
    nt, nm = struct.unpack_from("<II", raw, 8)
    p = 24
    for _ in range(nm):
        for _ in range(2):
            n, = struct.unpack_from("<H", raw, p)
            p += 2 + n
    return p


def test_tensor_entries_are_validated(good, tmp_path):
    raw = open(good, "rb").read()
    p0 = _tensor_table_pos(raw)
    nlen, = struct.unpack_from("<H", raw, p0)
    dt_pos = p0 + 2 + nlen                      # dtype u8, ndim u8, dims u32 x 2, offset u64, nbytes u64
    dims_pos, off_pos, nb_pos = dt_pos + 2, dt_pos + 2 + 8, dt_pos + 2 + 8 + 8
    p = _mutate(good, str(tmp_path / "dt.lqw"), lambda r: r.__setitem__(dt_pos, 7))
    assert "unknown dtype" in engine.check_model_file(p)
    p = _mutate(good, str(tmp_path / "nd.lqw"), lambda r: r.__setitem__(dt_pos + 1, 200))
    assert "rank" in engine.check_model_file(p)
    p = _mutate(good, str(tmp_path / "dim.lqw"), lambda r: r.__setitem__(slice(dims_pos, dims_pos + 4), struct.pack("<I", 3000)))
    assert "does not match its dims" in engine.check_model_file(p)
    p = _mutate(good, str(tmp_path / "nb.lqw"), lambda r: r.__setitem__(slice(nb_pos, nb_pos + 8), struct.pack("<Q", 1 << 50)))
    assert "does not match its dims" in engine.check_model_file(p)
    p = _mutate(good, str(tmp_path / "off.lqw"), lambda r: r.__setitem__(slice(off_pos, off_pos + 8), struct.pack("<Q", 1 << 30)))
    assert "outside the data section" in engine.check_model_file(p)
    p = _mutate(good, str(tmp_path / "mis.lqw"), lambda r: r.__setitem__(slice(off_pos, off_pos + 8), struct.pack("<Q", 8)))
    assert "misaligned" in engine.check_model_file(p)
    # truncated data section
    open(str(tmp_path / "tr.lqw"), "wb").write(raw[:-40])
    assert "outside the data section" in engine.check_model_file(str(tmp_path / "tr.lqw"))


def test_header_parser_survives_fuzzing_under_asan(good, tmp_path):
    """csrc/lqw_loader.h compiled into a small driver with -fsanitize=address,undefined (tests/native/lqw_fuzz.cpp): 20 000 mutated copies of a
    well-formed file (random header bytes, extreme 16/32/64-bit fields, truncation, shuffled bytes) -- every one is either accepted or
    rejected with a reason; an out-of-bounds read, an overflow or a crash aborts the driver."""
    import shutil
    import subprocess
    cuda_inc = "/usr/local/cuda/include"
    if shutil.which("g++") is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("g++ / CUDA headers not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "lqw_fuzz")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-I" + cuda_inc,
                        "-I" + os.path.join(root, "leaxer-qwen3-tts_b200", "csrc"), "-o", exe, os.path.join(root, "tests", "native", "lqw_fuzz.cpp")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 and "sanitize" in r.stdout:
        pytest.skip("sanitizer runtime not available: " + r.stdout[-200:])
    assert r.returncode == 0, r.stdout[-2000:]
    for seed in (1, 2):
        r = subprocess.run([exe, good, str(tmp_path / "scratch.lqw"), "10000", str(seed)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                           timeout=600, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
        assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout[-3000:]
        accepted, rejected = (int(x) for x in r.stdout.split()[1:3])
        assert accepted + rejected == 10000 and rejected > 4000      # about half are caught; the rest are harmless (padding, a flipped character in a name or meta value)
