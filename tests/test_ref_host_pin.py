"""Pins the oracle's HOST logic against the reference itself.

The reference's src/tts_onnx.cpp is compiled UNMODIFIED against oracle/ort_shim/onnxruntime_cxx_api.h (the real ONNX
Runtime is absent offline), whose Session::Run dispatches to deterministic stub graphs and logs every call. The Python
restatement (oracle/qwen3_tts_oracle.py: build_prompt_embeddings, generate_codes, predict_subcodes, sampler, vocoder
hand-off, clone front-end) is run over the numpy mirror of the same stubs and must reproduce the trace line for line:
same graphs in the same order (48 per frame), same tensor names and shapes, bit-identical inputs (prompt rows for
P = 8/9/10, all-ones masks, the KV tensors that go in and come back, the trailing-text / tts_pad schedule, the flattened
codes), same tokens with --top-k 1, same early stop on CODEC_EOS, same PCM length. The sampler filters are the reference's
own statics (exposed through `#define private public` in the driver's translation unit) and are compared bit for bit.

Where the binary is absent (no /root/reference and no prebuilt oracle/_ref) the same checks run against
tests/golden/ref_host_golden.npz, written from the reference build by tests/golden/make_ref_host_golden.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_host_cases as rc  # noqa: E402
from oracle import qwen3_tts_oracle as orc  # noqa: E402
from oracle import stub_graphs as sg  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "ref_host_golden.npz")
HAVE_SRC = os.path.exists("/root/reference/src/tts_onnx.cpp")


@pytest.fixture(scope="module")
def ref_bin():
    if HAVE_SRC:
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True, stderr=subprocess.DEVNULL)
    return rc.HOST_REF if os.path.exists(rc.HOST_REF) else None


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD, allow_pickle=False))


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("ref_host"))
    return d, rc.make_model_dir(d)


def _params(top_k, max_new):
    return orc.SamplingParams(temperature=0.8, top_k=top_k, top_p=0.95, max_new_tokens=max_new)


def _python_run(ids, lang, top_k, max_new, eos_at, speaker_embed=None, model=None):
    m = model or sg.StubModel(eos_at=eos_at)
    audio, codes = orc.synthesize_tokens(m, ids, lang, _params(top_k, max_new), speaker_embed=speaker_embed,
                                         schedule="reference")
    return m.trace, rc.result_line(np.asarray(audio, np.float32), m.graph_calls), codes


def _compare(name, trace, result, ref_bin, gold, ref_args):
    if ref_bin:
        ref_trace, ref_result = rc.run_ref(ref_args)
        assert len(trace) == len(ref_trace), (name, len(trace), len(ref_trace))
        for i, (a, b) in enumerate(zip(trace, ref_trace)):
            assert a == b, f"{name}: call {i} differs\n  oracle   : {a[:300]}\n  reference: {b[:300]}"
        assert result == ref_result, (name, result, ref_result)
    assert rc.sha(trace) == str(gold[f"{name}_sha"]), f"{name}: trace differs from the reference-made fixture"
    assert result == str(gold[f"{name}_result"])


@pytest.mark.parametrize("name", sorted(rc.ID_CASES))
def test_synthesize_tokens_trace(name, ref_bin, gold, workdir):
    lang, top_k, max_new, eos_at, text = rc.ID_CASES[name]
    ids = rc.wrap(text)
    assert ids == orc.wrap_text_ids(text)                                   # :243-259
    trace, result, codes = _python_run(ids, lang, top_k, max_new, eos_at)
    expect_P = 8 + (lang != "auto")
    assert f"inputs_embeds:1x{expect_P}x1024" in trace[[t.split()[0] for t in trace].index("talker_prefill")]
    frames = codes.shape[0]
    if eos_at > 0:
        assert frames == eos_at - expect_P < max_new                          # stopped by CODEC_EOS (:812)
    else:
        assert frames == max_new
    n_prompt_calls = 4 + max(0, len(text) - 1)                                # tts x3, codec batch, role, first text, trailing
    assert len(trace) == n_prompt_calls + 1 + 48 * frames + 1                 # prompt, prefill, 48 per frame (:801-846), vocoder
    _compare(name, trace, result, ref_bin, gold, ["ids", workdir[1], lang, 0.8, top_k, 0.95, max_new, eos_at, *ids])


def test_synthesize_text_through_reference_tokenizer(ref_bin, gold, workdir):
    """public synthesize(text): the reference tokenises with its own BPE (synthetic vocab next to the model dir)"""
    (name, (lang, top_k, max_new, eos_at, text)), = rc.TEXT_CASES.items()
    io_gold = np.load(os.path.join(ROOT, "tests", "golden", "io_reference.npz"))
    tok = io_gold[f"tok_{rc.io_cases.TOKENIZER_TEXTS.index(text)}"]
    assert tok[0] == 1
    ids = orc.wrap_text_ids([int(t) for t in tok[1:]])
    trace, result, _ = _python_run(ids, lang, top_k, max_new, eos_at)
    _compare(name, trace, result, ref_bin, gold, ["text", workdir[1], lang, 0.8, top_k, 0.95, max_new, eos_at, text])


def test_synthesize_clone_trace(ref_bin, gold, workdir):
    """public synthesize_clone(text, wav): WAV -> resample -> log-mel -> transposed speaker-encoder input -> prompt row (P = 10)"""
    (name, (lang, top_k, max_new, eos_at, text)), = rc.CLONE_CASES.items()
    io_gold = np.load(os.path.join(ROOT, "tests", "golden", "io_reference.npz"))
    tok = io_gold[f"tok_{rc.io_cases.TOKENIZER_TEXTS.index(text)}"]
    ids = orc.wrap_text_ids([int(t) for t in tok[1:]])
    wav = rc.write_ref_wav(os.path.join(workdir[0], "ref3s.wav"))
    mel = rc.ref_melwav(rc.IO_REF, wav) if os.path.exists(rc.IO_REF) else gold["clone_mel"]
    assert mel.shape == (128, 278)                                            # 3 s at hop 256, no centre padding
    if "clone_mel" in gold:
        assert np.array_equal(mel, gold["clone_mel"])
    m = sg.StubModel(eos_at=eos_at)
    spk = orc.extract_speaker_embedding(m, mel)
    trace, result, _ = _python_run(ids, lang, top_k, max_new, eos_at, speaker_embed=spk, model=m)
    assert trace[0].startswith("speaker_encoder mel:1x278x128")
    assert "inputs_embeds:1x10x1024" in trace[[t.split()[0] for t in trace].index("talker_prefill")]
    _compare(name, trace, result, ref_bin, gold, ["clone", workdir[1], lang, 0.8, top_k, 0.95, max_new, eos_at, wav, text])


def _ulp_diff(a, b):
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    return int(np.abs(ia - ib).max()) if a.size else 0


@pytest.mark.parametrize("V,k,p,seed", rc.FILTER_CASES)
def test_sampler_filters_bit_for_bit(V, k, p, seed, ref_bin, gold, tmp_path):
    """top_k_filter / softmax / top_p_filter (src/tts_onnx.cpp:907-950) executed by the reference vs the oracle's
    sampler_filtered_probs. The composition order and the renormalisation (:882-898) are checked through sample_token
    itself in test_sample_token_support."""
    x = rc.filter_logits(V, seed, ties=(seed % 3 == 2))
    key = f"filt_{V}_{k}_{seed}"
    if ref_bin:
        fin, fout = str(tmp_path / "in.f32"), str(tmp_path / "out.f32")
        x.tofile(fin)
        subprocess.run([ref_bin, "filt", fin, fout, str(k), str(p)], check=True)
        a = np.fromfile(fout, np.float32)[:V]                                 # top_k_filter(x)
        a.tofile(fin)
        subprocess.run([ref_bin, "filt", fin, fout, str(k), str(p)], check=True)
        o = np.fromfile(fout, np.float32)
        ref = np.stack([a, o[V:2 * V], o[2 * V:]])                            # top-k'd logits, softmax of them, top-p of that
        assert np.array_equal(ref, gold[key], equal_nan=True)
    ref = gold[key]
    a, sm, tp = ref[0], ref[1], ref[2]
    # (1) top-k: same survivors (ties at the threshold all survive: `x < threshold`, :923-926)
    only_k = orc.sampler_filtered_probs(x, orc.SamplingParams(temperature=1.0, top_k=k, top_p=1.0))
    assert np.array_equal(only_k > 0, np.isfinite(a))
    # (2) softmax: f32, serial sum in index order; the reference's expf vs the oracle's exp-in-f64: at most 1 ulp apart
    assert _ulp_diff(only_k[only_k > 0], sm[only_k > 0]) <= 1
    assert np.all(sm[only_k == 0] == 0)
    # (3) top-p on the reference's own probabilities + the renormalisation of :893-898, against the oracle end to end
    s2 = np.float32(0)
    for v in tp:
        s2 = np.float32(s2 + v)
    renorm = (tp / s2).astype(np.float32) if s2 > 0 else tp
    full = orc.sampler_filtered_probs(x, orc.SamplingParams(temperature=1.0, top_k=k, top_p=p))
    if np.array_equal(full > 0, renorm > 0):
        assert _ulp_diff(full[full > 0], renorm[full > 0]) <= 2
    else:
        # Exact ties straddling the cut: the reference's std::sort leaves the order of equal probabilities unspecified
        # (:934-935); the oracle and the kernels fix it as "lower index first". The cut itself (how many
        # survive, and with which probabilities) must still agree, and the disagreement must be among tied values only.
        a_s, b_s = np.sort(full[full > 0]), np.sort(renorm[renorm > 0])
        assert a_s.shape == b_s.shape and _ulp_diff(a_s, b_s) <= 2, "different top-p cutoff"
        moved = (full > 0) != (renorm > 0)
        assert len(set(sm[moved].tolist())) == 1, "survivor sets differ beyond an exact tie"


def test_sample_token_support(ref_bin, workdir, tmp_path):
    """sample_token (:878-905) as a whole: every token the reference draws lies in the support of the oracle's filtered
    distribution (temperature -> top-k -> softmax -> top-p -> renormalise), every likely token shows up, and top-k 1 or a
    tiny top-p make the unseeded draw deterministic = argmax."""
    if not ref_bin:
        pytest.skip("reference binary not built (needs /root/reference or a prebuilt oracle/_ref)")
    x = rc.filter_logits(2048, 11)
    fin = str(tmp_path / "in.f32")
    x.tofile(fin)

    def draw(temp, k, p, n):
        out = subprocess.run([ref_bin, "draw", workdir[1], fin, str(temp), str(k), str(p), str(n)], check=True,
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode().split("\n")
        line = [ln for ln in out if ln.startswith("TOKENS")][0]
        return np.asarray([int(t) for t in line.split()[1:]])
    for temp, k, p in [(0.8, 50, 0.95), (1.3, 8, 0.6), (0.5, 0, 0.9), (1.0, 50, 1.0)]:
        toks = draw(temp, k, p, 4000)
        prob = orc.sampler_filtered_probs(x, orc.SamplingParams(temperature=temp, top_k=k, top_p=p))
        assert np.all(prob[toks] > 0), (temp, k, p, "drawn outside the oracle's support")
        seen = np.bincount(toks, minlength=2048) / toks.size
        assert np.all(seen[prob > 0.02] > 0), "a likely token never drawn"
        assert np.abs(seen - prob).max() < 0.04
    assert np.all(draw(0.8, 1, 0.95, 50) == int(np.argmax(x)))
    assert np.all(draw(0.8, 50, 1e-4, 50) == int(np.argmax(x)))
