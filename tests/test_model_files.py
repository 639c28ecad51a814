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
