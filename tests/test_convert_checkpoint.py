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


@pytest.mark.gpu
def test_ort_cross_check(tmp_path):
    """Closes "parity unpinned" where the ingredients exist (none do offline, so this skips here): $LQT_ONNX_DIR = the reference's
    model directory with the seven .onnx graphs (README.md:71-93), $LQT_HF_CHECKPOINT = the HuggingFace checkpoint they were exported
    from, and the onnxruntime package. The checkpoint is converted (tools/convert_checkpoint.py), and the engine is compared with ONNX
    Runtime at the Ort::Session::Run boundary of src/tts_onnx.cpp: embeddings bit-for-bit in bf16-rounded weights,
    talker prefill logits and code-predictor logits within the 2e-2 bound, the vocoder within 1e-3 relative L2."""
    ort = pytest.importorskip("onnxruntime")
    onnx_dir, ckpt = os.environ.get("LQT_ONNX_DIR"), os.environ.get("LQT_HF_CHECKPOINT")
    if not onnx_dir or not ckpt:
        pytest.skip("set LQT_ONNX_DIR (reference .onnx graphs) and LQT_HF_CHECKPOINT (safetensors) to run the ORT cross-check")
    from __graft_entry__ import load_package
    load_package()
    from leaxer_qwen3_tts_b200 import engine
    spec = ms.spec_0p6b(0)
    out = str(tmp_path / "onnx_kv")
    problems = cc.convert(cc.load_safetensors_dir(ckpt), spec, out)
    assert problems == [], problems[:5]
    eng = engine.Engine(out, device=0, kv_dtype="f32")

    def sess(name):
        return ort.InferenceSession(os.path.join(onnx_dir, name + ".onnx"), providers=["CPUExecutionProvider"])

    rng = np.random.default_rng(0)
    # text_project / codec_embed / code_predictor_embed  (src/tts_onnx.cpp:545-613)
    ids = rng.integers(0, 151643, size=(1, 12)).astype(np.int64)
    ref = sess("text_project").run(None, {"input_ids": ids})[0][0]
    assert np.abs(eng.text_project(ids[0]) - ref).max() < 2e-2
    cids = rng.integers(0, spec.vocab, size=(1, 8)).astype(np.int64)
    ref = sess("codec_embed").run(None, {"input_ids": cids})[0][0]
    assert np.abs(eng.codec_embed(cids[0]) - ref).max() < 1e-2
    ref = sess("code_predictor_embed").run(None, {"input_ids": np.array([[77]], np.int64), "generation_step": np.array([3], np.int64)})[0]
    assert np.abs(eng.code_predictor_embed(77, 3) - ref.reshape(-1)).max() < 1e-2
    # talker prefill: logits of the last row + last_hidden (:615-665)
    emb = (rng.standard_normal((1, 9, spec.hidden)) * 0.5).astype(np.float32)
    outs = sess("talker_prefill").run(["logits", "last_hidden"], {"inputs_embeds": emb, "attention_mask": np.ones((1, 9), np.int64)})
    logits, hidden = eng.talker_prefill(emb[0])
    assert np.abs(logits - outs[0][0, -1]).max() < 2e-2
    assert np.abs(hidden - outs[1].reshape(-1)[: spec.hidden]).max() < 2e-2
    # code predictor (:734-757): head `generation_step` on the last of L rows
    rows = (rng.standard_normal((1, 5, spec.hidden)) * 0.5).astype(np.float32)
    ref = sess("code_predictor").run(None, {"inputs_embeds": rows, "generation_step": np.array([3], np.int64)})[0].reshape(-1)[: spec.cp_vocab]
    assert np.abs(eng.code_predictor(rows[0], 3) - ref).max() < 2e-2
    # vocoder (:759-776)
    codes = rng.integers(0, 2048, size=(1, 40, 16)).astype(np.int64)
    outs = sess("tokenizer12hz_decode").run(None, {"audio_codes": codes})
    n = int(np.asarray(outs[1]).reshape(-1)[0])
    ref = np.asarray(outs[0]).reshape(-1)[:n]
    got = eng.vocoder_decode(codes[0])
    assert got.shape[0] == n == 40 * 1920
    assert np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-12) < 1e-3
    eng.close()
