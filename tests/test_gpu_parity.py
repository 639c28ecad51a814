"""GPU parity tests: every C-ABI entry point against the CPU oracle on the same seeded inputs.
Tolerances (BASELINE.json north_star): token ids bit-exact (greedy and seeded sampler), logits
<= 2e-2 max-abs, waveform <= 1e-3 relative L2 and >= 40 dB SNR. The batch-1 path keeps fp32
activations, so the observed errors are orders of magnitude below those bounds; the tests assert
the north_star bound and a tighter engineering bound."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LOGIT_TOL = 2e-2          # north_star
LOGIT_TOL_TIGHT = 2e-3    # engineering bound for the fp32-activation path
WAVE_REL_L2 = 1e-3
WAVE_SNR_DB = 40.0


def rnd(shape, seed, scale=1.0):
    g = np.random.default_rng(seed)
    return (g.standard_normal(shape) * scale).astype(np.float32)


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def snr_db(a, ref):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return float(10 * np.log10((ref ** 2).sum() / (((a - ref) ** 2).sum() + 1e-30)))


def pair(request, which):
    return request.getfixturevalue(f"{which}_engine"), request.getfixturevalue(f"{which}_oracle")


def tight(which):
    """Engineering bound. The default engine rounds talker K/V to bf16 when they enter the paged
    cache (north_star); a value that lands within fp32 summation noise of a bf16 rounding boundary
    can round the other way than in the oracle, which on the full-size model moves logits by up to
    ~1e-2 (still inside the north_star 2e-2). The fp32-KV parity mode ("full_f32") has no such
    rounding point and must meet the tight bound."""
    return LOGIT_TOL if which == "full" else LOGIT_TOL_TIGHT


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("which", ["tiny", "full"])
def test_embeddings(request, which):
    eng, m = pair(request, which)
    ids = [151672, 151673, 151671, 151644, 77091, 0, 1, 14990, 151642, 99999, 5]
    out = eng.text_project(ids)
    ref = m.text_project(ids).numpy()
    assert out.shape == ref.shape
    assert rel_l2(out, ref) < 1e-5, rel_l2(out, ref)
    cids = [2148, 2149, 2150, 2154, 2155, 2156, 2157, 2050, 0, 2047, 3071]
    assert np.array_equal(eng.codec_embed(cids), m.codec_embed(cids).numpy())
    for step, tok in [(0, 0), (3, 1234), (14, 2047)]:
        assert np.array_equal(eng.code_predictor_embed(tok, step), m.code_predictor_embed(tok, step).numpy())


@pytest.mark.parametrize("which", ["tiny", "full", "full_f32"])
def test_talker_prefill_and_decode(request, which):
    eng, m = pair(request, which)
    H = m.spec.hidden
    P = 9
    x = rnd((P, H), 1)
    kv = m.new_kv()
    ref_logits, ref_hid = m.talker_prefill(torch.from_numpy(x), kv)
    logits, hid = eng.talker_prefill(x, slot=0)
    assert eng.kv_len(0) == P
    e1, e2 = maxabs(logits, ref_logits[-1].numpy()), maxabs(hid, ref_hid.numpy())
    assert e1 < LOGIT_TOL and e1 < tight(which), e1
    assert e2 < tight(which), e2
    for step in range(4):
        e = rnd((H,), 10 + step, 4.0)
        rl, rh = m.talker_decode(torch.from_numpy(e), kv)
        lg, hd = eng.talker_decode(e, slot=0)
        e1, e2 = maxabs(lg, rl.numpy()), maxabs(hd, rh.numpy())
        assert e1 < tight(which) and e2 < tight(which), (step, e1, e2)
        assert int(np.argmax(lg)) == int(np.argmax(rl.numpy()))
    assert eng.kv_len(0) == P + 4


def test_talker_long_context_split_kv(request):
    """crosses several 64-position KV pages so that every attention split and the combine run"""
    eng, m = pair(request, "tiny")
    H = m.spec.hidden
    n = 150
    kv = m.new_kv()
    x = rnd((n, H), 3)
    m.talker_prefill(torch.from_numpy(x[:1]), kv)
    eng.talker_prefill(x[:1], slot=1)
    worst = 0.0
    for i in range(1, n):
        rl, rh = m.talker_decode(torch.from_numpy(x[i]), kv)
        lg, hd = eng.talker_decode(x[i], slot=1)
        worst = max(worst, maxabs(lg, rl.numpy()), maxabs(hd, rh.numpy()))
    assert worst < LOGIT_TOL_TIGHT, worst


@pytest.mark.parametrize("which", ["full_f32", "full"])
def test_talker_context_beyond_one_kv_round(request, which):
    """More than 480 positions: every attention split holds more than 32 positions, so a warp runs several rounds of four cached K/V
    rows (the first round is requested before the grid hand-over, the later ones inside the loop; frame_kernel.cuh
    talker_attn_partial), the last round is ragged, and 8 KV pages are addressed through the page-table copy in shared memory."""
    eng, m = pair(request, which)
    H = m.spec.hidden
    P = 500
    x = rnd((P + 3, H), 17)
    kv = m.new_kv()
    ref_logits, ref_hid = m.talker_prefill(torch.from_numpy(x[:P]), kv)
    logits, hid = eng.talker_prefill(x[:P], slot=0)
    assert eng.kv_len(0) == P
    assert maxabs(logits, ref_logits[-1].numpy()) < tight(which) and maxabs(hid, ref_hid.numpy()) < tight(which)
    for i in range(P, P + 3):
        rl, rh = m.talker_decode(torch.from_numpy(x[i]), kv)
        lg, hd = eng.talker_decode(x[i], slot=0)
        e1, e2 = maxabs(lg, rl.numpy()), maxabs(hd, rh.numpy())
        assert e1 < tight(which) and e2 < tight(which), (i, e1, e2)
    eng.kv_reset(0)


@pytest.mark.parametrize("which", ["tiny", "full"])
def test_code_predictor(request, which):
    eng, m = pair(request, which)
    H = m.spec.hidden
    for L, step in [(2, 0), (3, 1), (9, 7), (16, 14)]:
        x = rnd((L, H), 100 + L)
        ref = m.code_predictor(torch.from_numpy(x), step).numpy()
        out = eng.code_predictor(x, step)
        e = maxabs(out, ref)
        assert e < LOGIT_TOL_TIGHT, (L, step, e)
        assert int(np.argmax(out)) == int(np.argmax(ref))


def test_sampler_bit_exact(request):
    """token-exact given matching logits, over filters, ties, masks and the Philox stream"""
    eng, m = pair(request, "tiny")
    orc = request.getfixturevalue("oracle_mod")
    cases = []
    for V in (2048, 3072):
        for i, (t, k, p) in enumerate([(0.8, 50, 0.95), (1.0, 50, 0.95), (0.0, 50, 0.95), (0.8, 0, 0.95),
                                       (0.8, 50, 1.0), (1.3, 5, 0.5), (0.8, 1, 0.95), (0.7, 4000, 0.9),
                                       (0.9, 0, 1.0)]):
            cases.append((V, t, k, p, i))
    n_checked = 0
    for V, t, k, p, i in cases:
        for rep in range(6):
            lg = rnd((V,), 1000 * i + rep + V, 3.2)
            if rep == 1:                       # ties at the top-k threshold and among the top probs
                lg = np.round(lg * 2) / 2
            if rep == 2:
                lg[:7] = lg.max() + 1.0        # exact ties at the maximum
            mask = (V == 3072)
            frame, cb = 7 * rep + i, (rep * 5 + i) % 16
            sp = orc.SamplingParams(temperature=t, top_k=k, top_p=p, seed=1234 + i, utterance_id=rep)
            lref = lg.copy()
            if mask:
                keep = lref[2150]
                lref[2048:] = -np.inf
                lref[2150] = keep
            ref = orc.sample_token(lref, sp, frame, cb)
            got = eng.sample(lg, eng.sampling(t, k, p, 1, 1234 + i, rep), frame, cb, mask_codec_specials=mask)
            assert got == ref, (V, t, k, p, rep, got, ref)
            n_checked += 1
        # greedy: lowest index among exact ties
        lg = np.round(rnd((V,), 77 + i, 3.0))
        sp = orc.SamplingParams(greedy=True)
        assert eng.sample(lg, eng.sampling(greedy=True), 0, 0) == orc.sample_token(lg, sp, 0, 0)
    assert n_checked >= 100


@pytest.mark.parametrize("which,lang,with_spk", [("tiny", "auto", False), ("tiny", "en", False),
                                                 ("tiny", "ko", True), ("tiny", "auto", True),
                                                 ("full", "zh", True)])
def test_build_prompt(request, which, lang, with_spk):
    eng, m = pair(request, which)
    orc = request.getfixturevalue("oracle_mod")
    ids = orc.wrap_text_ids([14990, 14615, 88225, 20339, 13189])
    spk = rnd((m.spec.hidden,), 5) if with_spk else None
    st = orc.UtteranceState(kv=m.new_kv())
    ref = orc.build_prompt_embeddings(m, ids, lang, st, spk).numpy()
    prompt, trailing, pad = eng.build_prompt(ids, lang, spk)
    expect_P = 8 + (1 if lang != "auto" else 0) + (1 if with_spk else 0)
    assert prompt.shape == ref.shape == (expect_P, m.spec.hidden)
    assert rel_l2(prompt, ref) < 1e-5
    assert trailing.shape[0] == st.trailing_len == 5
    assert rel_l2(trailing, st.trailing_text_hidden.numpy()) < 1e-5
    assert rel_l2(pad, st.tts_pad_embed.numpy()) < 1e-5


@pytest.mark.parametrize("which,T", [("tiny", 7), ("full", 12)])
def test_vocoder(request, which, T):
    eng, m = pair(request, which)
    codes = np.random.default_rng(T).integers(0, 2048, size=(T, 16))
    ref, n = m.vocoder(codes)
    ref = ref.numpy()
    out = eng.vocoder_decode(codes)
    assert out.shape[0] == n == T * m.spec.samples_per_frame
    assert rel_l2(out, ref) < WAVE_REL_L2, rel_l2(out, ref)
    assert snr_db(out, ref) > WAVE_SNR_DB
    assert np.abs(out).max() <= 1.0


@pytest.mark.parametrize("which", ["tiny", "full"])
def test_speaker_encoder(request, which):
    eng, m = pair(request, which)
    mel = rnd((278, 128), 9, 2.0) - 5.0
    ref = m.speaker_encoder(torch.from_numpy(mel)).numpy()
    out = eng.speaker_encoder(mel)
    assert rel_l2(out, ref) < 1e-4, rel_l2(out, ref)


def _run_generate(eng, m, orc, ids, lang, sp_o, sp_e, trace=True):
    st = orc.UtteranceState(kv=m.new_kv())
    prompt = orc.build_prompt_embeddings(m, ids, lang, st)
    tr = {}
    ref_codes = np.asarray(orc.generate_codes(m, prompt, st, sp_o, trace=tr), dtype=np.int64).reshape(-1, 16)
    got = eng.generate(prompt.numpy(), st.trailing_text_hidden.numpy(), st.tts_pad_embed.numpy(), sp_e,
                       trace=trace)
    return ref_codes, tr, got


@pytest.mark.parametrize("which,frames", [("tiny", 12), ("full", 6), ("full_f32", 6)])
def test_generate_greedy_token_exact(request, which, frames):
    eng, m = pair(request, which)
    orc = request.getfixturevalue("oracle_mod")
    ids = orc.wrap_text_ids([14990, 14615, 88225])
    sp_o = orc.SamplingParams(max_new_tokens=frames, greedy=True)
    ref_codes, tr, (codes, tb) = _run_generate(eng, m, orc, ids, "en", sp_o, eng.sampling(max_new_tokens=frames, greedy=True))
    assert codes.shape == ref_codes.shape == (frames, 16)
    V, Vs = m.spec.vocab, m.spec.cp_vocab
    # Free-running greedy decoding is token-exact UNTIL the oracle itself is at a near-tie: a top-2 margin below the
    # logit tolerance cannot be resolved by any implementation that sums in a different order ("exact-token
    # parity is numerically fragile"). A flip is accepted only there, only towards the oracle's runner-up, and ends the comparison.
    diff = np.argwhere(codes != ref_codes)
    n_cmp = frames * 16 if len(diff) == 0 else int(diff[0][0]) * 16 + int(diff[0][1])
    if len(diff):
        f, j = int(diff[0][0]), int(diff[0][1])
        r = np.asarray(tr["talker_logits"][f] if j == 0 else tr["cp_logits"][f][j - 1], dtype=np.float64)
        order = np.argsort(-np.where(np.isfinite(r), r, -np.inf))
        margin = r[order[0]] - r[order[1]]
        assert margin < LOGIT_TOL_TIGHT, (f, j, margin, "token flip away from a near-tie")
        assert int(codes[f, j]) == int(order[1]) and int(ref_codes[f, j]) == int(order[0]), (f, j, codes[f, j], order[:2])
        assert n_cmp >= 16, "diverged inside the first frame"
    for f in range(frames):
        for j in range(16):
            if f * 16 + j > n_cmp:
                break
            ref = tr["talker_logits"][f] if j == 0 else tr["cp_logits"][f][j - 1]
            got = tb[f, 0, :V] if j == 0 else tb[f, j, :Vs]
            fin = np.isfinite(ref)
            assert maxabs(got[fin], np.asarray(ref)[fin]) < tight(which), (f, j)
            assert np.all(np.isneginf(got[~fin]))


@pytest.mark.parametrize("which,frames", [("tiny", 12), ("full", 5)])
def test_generate_seeded_sampler_token_exact(request, which, frames):
    eng, m = pair(request, which)
    orc = request.getfixturevalue("oracle_mod")
    ids = orc.wrap_text_ids([1000, 2000, 3000, 4000, 5000, 6000])
    sp_o = orc.SamplingParams(temperature=0.8, top_k=50, top_p=0.95, max_new_tokens=frames, seed=1234, utterance_id=3)
    sp_e = eng.sampling(0.8, 50, 0.95, frames, 1234, 3)
    ref_codes, tr, (codes, tb) = _run_generate(eng, m, orc, ids, "auto", sp_o, sp_e)
    assert np.array_equal(codes, ref_codes), np.argwhere(codes != ref_codes)[:4]


@pytest.mark.parametrize("clusters", [13, 14])
def test_generate_other_grid_sizes(request, full_dir, monkeypatch, clusters):
    """The weight slices (whole 8-row tiles per CTA, tile pairs, K split per warp) depend on the number of co-resident
    clusters: 13 and 14 clusters give other tile counts, single tiles and K splits than the default 15 (fewer than 13
    clusters exceed the per-CTA row limits and fall back to the graph path). Seeded sampling
    must stay token-exact against the oracle, and the logits within tolerance."""
    from leaxer_qwen3_tts_b200 import engine
    m = request.getfixturevalue("full_oracle")
    orc = request.getfixturevalue("oracle_mod")
    monkeypatch.setenv("LQT_FK_CLUSTERS", str(clusters))
    eng = engine.Engine(full_dir, device=0)
    try:
        frames = 3
        ids = orc.wrap_text_ids([1000, 2000, 3000, 4000, 5000, 6000])
        sp_o = orc.SamplingParams(temperature=0.8, top_k=50, top_p=0.95, max_new_tokens=frames, seed=1234, utterance_id=3)
        eng.reset_stats()
        ref_codes, tr, (codes, tb) = _run_generate(eng, m, orc, ids, "auto", sp_o, eng.sampling(0.8, 50, 0.95, frames, 1234, 3))
        assert eng.stats().kernel_launches < 40, "the persistent kernel did not run (graph fallback?)"
        assert np.array_equal(codes, ref_codes), np.argwhere(codes != ref_codes)[:4]
        V, Vs = m.spec.vocab, m.spec.cp_vocab
        for f in range(frames):
            ref0 = np.asarray(tr["talker_logits"][f])
            fin = np.isfinite(ref0)
            assert maxabs(tb[f, 0, :V][fin], ref0[fin]) < LOGIT_TOL
            assert maxabs(tb[f, 1:, :Vs], tr["cp_logits"][f]) < LOGIT_TOL
    finally:
        eng.close()


def test_teacher_forced_logits(request):
    """feed the oracle's tokens; logits of every one of the 16 draws per frame stay within tolerance
    and the argmax agrees wherever the oracle's top-2 margin exceeds the tolerance"""
    eng, m = pair(request, "tiny")
    orc = request.getfixturevalue("oracle_mod")
    frames = 40
    ids = orc.wrap_text_ids(orc.synthetic_text_ids(20))
    sp_o = orc.SamplingParams(max_new_tokens=frames, seed=7)
    st = orc.UtteranceState(kv=m.new_kv())
    prompt = orc.build_prompt_embeddings(m, ids, "ja", st)
    tr = {}
    ref_codes = np.asarray(orc.generate_codes(m, prompt, st, sp_o, trace=tr), dtype=np.int64)
    sp_e = eng.sampling(max_new_tokens=frames, seed=99)           # different seed: tokens come from forcing
    codes, tb = eng.generate(prompt.numpy(), st.trailing_text_hidden.numpy(), st.tts_pad_embed.numpy(), sp_e,
                             forced_codes=ref_codes, trace=True)
    assert np.array_equal(codes, ref_codes)
    V, Vs = m.spec.vocab, m.spec.cp_vocab
    worst = 0.0
    for f in range(frames):
        ref0 = tr["talker_logits"][f]
        fin = np.isfinite(ref0)
        worst = max(worst, maxabs(tb[f, 0, :V][fin], ref0[fin]), maxabs(tb[f, 1:, :Vs], tr["cp_logits"][f]))
        for j in range(15):
            r = tr["cp_logits"][f][j]
            top2 = np.sort(r)[-2:]
            if top2[1] - top2[0] > LOGIT_TOL:
                assert int(np.argmax(tb[f, 1 + j, :Vs])) == int(np.argmax(r))
    assert worst < LOGIT_TOL and worst < LOGIT_TOL_TIGHT, worst


def test_eos_and_limits(request):
    eng, m = pair(request, "tiny")
    orc = request.getfixturevalue("oracle_mod")
    ids = orc.wrap_text_ids([5, 6, 7])
    st = orc.UtteranceState(kv=m.new_kv())
    prompt = orc.build_prompt_embeddings(m, ids, "auto", st).numpy()
    tr, pad = st.trailing_text_hidden.numpy(), st.tts_pad_embed.numpy()
    # max_new_tokens = 0 -> no frames (loop A never runs, src/tts_onnx.cpp:801)
    assert eng.generate(prompt, tr, pad, eng.sampling(max_new_tokens=0, greedy=True)).shape == (0, 16)
    # forced EOS as the first code0 -> empty result (src/tts_onnx.cpp:812, 418)
    forced = np.zeros((3, 16), np.int64)
    forced[0, 0] = 2150
    assert eng.generate(prompt, tr, pad, eng.sampling(max_new_tokens=3, greedy=True), forced_codes=forced).shape == (0, 16)
    # EOS at frame 2 -> exactly 2 frames, and the engine is reusable afterwards
    forced = np.random.default_rng(0).integers(0, 2048, size=(5, 16))
    forced[2, 0] = 2150
    out = eng.generate(prompt, tr, pad, eng.sampling(max_new_tokens=5, greedy=True), forced_codes=forced)
    assert out.shape == (2, 16) and np.array_equal(out, forced[:2])
    again = eng.generate(prompt, tr, pad, eng.sampling(max_new_tokens=2, greedy=True))
    assert again.shape == (2, 16)
    # error paths: too few ids, out-of-range ids (no crash, error string set)
    from leaxer_qwen3_tts_b200.engine import EngineError
    with pytest.raises(EngineError):
        eng.build_prompt([1, 2, 3, 4])
    with pytest.raises(EngineError):
        eng.text_project([151936])
    with pytest.raises(EngineError):
        eng.vocoder_decode(np.full((2, 16), 2048))


def test_synthesize_tokens_matches_golden(request):
    """C1 (BASELINE.json configs[0]): 'Hello world' en greedy through lqt_synthesize_tokens vs the
    committed golden fixture (tests/golden/make_golden.py)."""
    import os
    eng, m = pair(request, "full")
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "c1_hello_world_en_greedy.npz"))
    audio, codes = eng.synthesize_tokens(g["token_ids"], "en", max_new_tokens=int(g["codes"].shape[0]), greedy=True)
    assert np.array_equal(codes, g["codes"])
    assert rel_l2(audio, g["audio"]) < WAVE_REL_L2
    assert snr_db(audio, g["audio"]) > WAVE_SNR_DB
    # size-independent property: the vocoder is causal, so a prefix of the codes gives a prefix of the audio
    half = eng.vocoder_decode(g["codes"][:10])
    assert rel_l2(half, audio[: half.shape[0]]) < 1e-5


def test_first_audio_chunk_is_exact_prefix(tiny_dir, monkeypatch):
    """the first chunk (25 frames) is vocoded early on a second stream; every vocoder op is causal, so the result
    must be bit-identical to a run without chunking, and the first-audio time must be shorter than the full latency."""
    from leaxer_qwen3_tts_b200 import engine
    ids = engine.wrap_text_ids([14990, 14615, 88225, 20339])
    monkeypatch.setenv("LQT_FIRST_CHUNK", "0")
    e0 = engine.Engine(tiny_dir, device=0)
    a0, c0 = e0.synthesize_tokens(ids, "en", max_new_tokens=60, seed=5, utterance_id=1)
    s0 = e0.stats()
    assert abs(s0.first_audio_ms - s0.last_total_ms) < 1e-3
    a0s, c0s = e0.synthesize_tokens(ids, "en", max_new_tokens=12, seed=5, utterance_id=1)
    e0.close()
    monkeypatch.delenv("LQT_FIRST_CHUNK")                 # defaults: 4 frames first, then 25 -- also for utterances shorter than 25
    e2 = engine.Engine(tiny_dir, device=0)
    a2s, c2s = e2.synthesize_tokens(ids, "en", max_new_tokens=12, seed=5, utterance_id=1)
    assert np.array_equal(c0s, c2s) and np.array_equal(a0s, a2s) and 0 < e2.stats().first_audio_ms < e2.stats().last_total_ms
    a2l, c2l = e2.synthesize_tokens(ids, "en", max_new_tokens=60, seed=5, utterance_id=1)
    assert np.array_equal(c0, c2l) and np.array_equal(a0, a2l)
    e2.close()
    monkeypatch.setenv("LQT_FIRST_CHUNK", "25")
    e1 = engine.Engine(tiny_dir, device=0)
    a1, c1 = e1.synthesize_tokens(ids, "en", max_new_tokens=60, seed=5, utterance_id=1)
    s1 = e1.stats()
    assert np.array_equal(c0, c1) and np.array_equal(a0, a1)
    assert 0 < s1.first_audio_ms < s1.last_total_ms
    a2, c2 = e1.synthesize_tokens(ids, "en", max_new_tokens=20, seed=5, utterance_id=1)      # shorter than the chunk: no split
    assert c2.shape[0] == 20 and np.array_equal(c2, c0[:20])
    e1.close()


# ---------------------------------------------------------------------------------------------
# Parity at the BENCHMARKED shape (BASELINE.json configs[1]: full 0.6B, bf16 KV, 375 frames = 384 positions,
# 6 KV pages, 12 attention splits), against tests/golden/c2_full_375.npz (oracle-made: tests/golden/make_c2_golden.py)
# ---------------------------------------------------------------------------------------------
def _c2_golden():
    import os
    from conftest import GOLDEN
    return np.load(os.path.join(GOLDEN, "c2_full_375.npz"))


def test_c2_teacher_forced_375_frames(request):
    """the engine is fed the oracle's tokens for the whole 30 s utterance (so that the two never diverge) and must
    reproduce the oracle's logits within the north_star bound at every stored frame, from the first to the last (KV pages 0-5,
    every attention split), with the same argmax wherever the oracle's own margin exceeds the bound"""
    eng, _ = pair(request, "full")
    g = _c2_golden()
    frames = int(g["codes"].shape[0])
    assert frames == 375
    prompt, trailing, pad = eng.build_prompt(g["token_ids"], "en")
    assert prompt.shape[0] == 9 and trailing.shape[0] == 90
    sp = eng.sampling(0.8, 50, 0.95, frames, seed=4321, utterance_id=9)     # another stream: tokens come from forcing
    codes, tb = eng.generate(prompt, trailing, pad, sp, forced_codes=g["codes"], trace=True)
    assert np.array_equal(codes, g["codes"])
    assert eng.kv_len(0) == 9 + frames
    V, Vs = 3072, 2048
    worst = 0.0
    for i, f in enumerate(g["frames"]):
        ref0, refc = g["talker_logits"][i], g["cp_logits"][i]
        fin = np.isfinite(ref0)
        e0, ec = maxabs(tb[f, 0, :V][fin], ref0[fin]), maxabs(tb[f, 1:, :Vs], refc)
        assert np.all(np.isneginf(tb[f, 0, :V][~fin]))
        worst = max(worst, e0, ec)
        assert e0 < LOGIT_TOL and ec < LOGIT_TOL, (int(f), e0, ec)
        for row_ref, row in [(ref0[fin], tb[f, 0, :V][fin])] + [(refc[j], tb[f, 1 + j, :Vs]) for j in range(15)]:
            top2 = np.sort(row_ref)[-2:]
            if top2[1] - top2[0] > LOGIT_TOL:
                assert int(np.argmax(row)) == int(np.argmax(row_ref)), int(f)
    print(f"\n[c2 teacher-forced] worst |logit error| over {len(g['frames'])} stored frames: {worst:.2e}")


@pytest.mark.parametrize("which,key,min_prefix", [("full", "codes", 64), ("full_f32", "codes_f32kv", 375)])   # measured on B200: 146 and 375
def test_c2_free_running_seeded_prefix(request, which, key, min_prefix):
    """free-running C2 with the benchmark's sampler settings and Philox key (1234, 0): token-exact against the oracle for as
    long as no draw sits within float noise of a CDF boundary. The fp32-KV parity mode has no rounding point and must
    sustain it far longer than the bf16-KV default (whose K/V rounding can flip on a boundary, moving logits by ~1e-2)."""
    eng, _ = pair(request, which)
    g = _c2_golden()
    ref = g[key]
    prompt, trailing, pad = eng.build_prompt(g["token_ids"], "en")
    codes = eng.generate(prompt, trailing, pad, eng.sampling(0.8, 50, 0.95, 375, seed=1234, utterance_id=0))
    assert codes.shape == ref.shape
    diff = np.argwhere(codes != ref)
    prefix = 375 if len(diff) == 0 else int(diff[0][0])
    print(f"\n[c2 free-running {which}] token-exact prefix: {prefix} of 375 frames"
          + ("" if len(diff) == 0 else f" (first difference: frame {diff[0][0]}, codebook {diff[0][1]})"))
    assert prefix >= min_prefix, (prefix, diff[:3])
    assert codes.min() >= 0 and codes[:, 0].max() < 2048 and codes.max() < 2048


@pytest.mark.parametrize("top_k,top_p,temp", [(0, 0.95, 0.8), (4000, 0.9, 0.7), (100, 1.0, 1.0), (50, 0.95, 0.0), (65, 0.5, 1.3)])
def test_generate_sampler_general_path(request, top_k, top_p, temp):
    """the IN-KERNEL sampler's general path (frame_kernel.cuh fk_sample: radix select, rank, top-p) runs whenever top_k is 0,
    larger than 64 or not below the vocabulary; the fast path covers 0 < top_k <= 64. Token-exact against the oracle through
    lqt_generate (not the standalone lqt_sample kernel)."""
    eng, m = pair(request, "tiny")
    orc = request.getfixturevalue("oracle_mod")
    frames = 10
    ids = orc.wrap_text_ids([777, 888, 999, 1111])
    sp_o = orc.SamplingParams(temperature=temp, top_k=top_k, top_p=top_p, max_new_tokens=frames, seed=99, utterance_id=5)
    sp_e = eng.sampling(temp, top_k, top_p, frames, 99, 5)
    ref_codes, tr, (codes, tb) = _run_generate(eng, m, orc, ids, "ko", sp_o, sp_e)
    assert np.array_equal(codes, ref_codes), np.argwhere(codes != ref_codes)[:4]


@pytest.mark.parametrize("which,T", [("tiny", 24), ("full", 100)])
def test_vocoder_beyond_the_attention_window(request, which, T):
    """T above the sliding window (8 tiny / 72 full) so that window_attn_kernel's mask takes effect, against the oracle"""
    eng, m = pair(request, which)
    assert T > m.spec.voc_window
    codes = np.random.default_rng(1000 + T).integers(0, 2048, size=(T, 16))
    ref, n = m.vocoder(codes)
    ref = ref.numpy()
    out = eng.vocoder_decode(codes)
    assert out.shape[0] == n == T * m.spec.samples_per_frame
    assert rel_l2(out, ref) < WAVE_REL_L2, rel_l2(out, ref)
    assert snr_db(out, ref) > WAVE_SNR_DB
    # the tail (frames beyond the window) separately: an unmasked attention would only show there
    tail = slice((m.spec.voc_window + 4) * m.spec.samples_per_frame, None)
    assert rel_l2(out[tail], ref[tail]) < WAVE_REL_L2


def test_create_rejects_wrong_shapes(tiny_dir, tmp_path):
    """a weight file whose tensor is smaller than the model spec says must fail lqt_create with a reason (ADVICE r1:
    engine.cu need<T>), not become an out-of-bounds device read"""
    import os
    import shutil
    from leaxer_qwen3_tts_b200 import engine
    bad = str(tmp_path / "onnx_kv")
    shutil.copytree(tiny_dir, bad)
    spec = ms_mod().spec_tiny(0)
    meta = dict(spec.to_meta()); meta["graph"] = "codec_embed"
    small = ms_mod().f32_to_bf16_bits(np.zeros((spec.vocab - 8, spec.hidden), np.float32))
    ms_mod().write_lqw(os.path.join(bad, "codec_embed.lqw"), [("embed", ms_mod().DT_BF16, small)], meta)
    with pytest.raises(engine.EngineError, match="tensor embed: expected bf16"):
        engine.Engine(bad, device=0)
    # wrong dtype
    f32 = np.zeros((spec.vocab, spec.hidden), np.float32)
    ms_mod().write_lqw(os.path.join(bad, "codec_embed.lqw"), [("embed", ms_mod().DT_F32, f32)], meta)
    with pytest.raises(engine.EngineError, match="expected bf16"):
        engine.Engine(bad, device=0)


def ms_mod():
    from leaxer_qwen3_tts_b200 import modelspec
    return modelspec


def test_strict_persistent_fails_loudly_and_auto_reports(request):
    """1.7B does not fit the persistent kernel: an explicit request fails with the reason; AUTO (what lqt_create uses) runs
    and says which frame loop is active"""
    from leaxer_qwen3_tts_b200 import engine, modelspec
    spec = modelspec.ModelSpec(name="qwen3-tts-wide-test", hidden=2048, inter=6144, layers=2, cp_layers=1, text_dim=64,
                               max_pos=64, voc_max_pos=64, voc_codebook_dim=32, voc_rvq_out=64, voc_hidden=128, voc_layers=1,
                               voc_heads=2, voc_inter=256, voc_window=8, voc_decoder_dim=192, spk_channels=64, seed=3)
    d = modelspec.generate_model_dir(modelspec.default_model_dir(spec), spec)
    with pytest.raises(engine.EngineError, match="persistent frame kernel unavailable"):
        engine.Engine(d, device=0, frame_impl="persistent")
    e = engine.Engine(d, device=0, frame_impl="auto")
    assert e.stats().frame_impl_active == engine.FRAME_IMPL["batched"]
    # ... and the plain single-utterance call works on such a handle (one slot of the batched tcgen05 path), token-exact
    orc = request.getfixturevalue("oracle_mod")
    ids = orc.wrap_text_ids([31, 32, 33])
    sp = orc.SamplingParams(temperature=0.8, top_k=50, top_p=0.95, max_new_tokens=5, seed=11, utterance_id=2)
    ref_audio, ref_codes = orc.synthesize_tokens(orc.OracleModel(d), ids, "en", sp)
    audio, codes = e.synthesize_tokens(ids, "en", 0.8, 50, 0.95, 5, 11, 2)
    assert np.array_equal(codes, ref_codes) and rel_l2(audio, ref_audio) < WAVE_REL_L2
    e.close()
    assert request.getfixturevalue("tiny_engine").stats().frame_impl_active == engine.FRAME_IMPL["persistent"]


def test_no_cooperative_launch_disables_the_overlap(tiny_dir, monkeypatch):
    """LQT_FK_NOCOOP (profilers that cannot replay cooperative launches; also the state after a refused launch): co-residency
    is then only assumed, so the chunk vocoder must not run beside the frame kernel (engine.cu:1050)"""
    from leaxer_qwen3_tts_b200 import engine
    ids = engine.wrap_text_ids([14990, 14615, 88225, 20339])
    monkeypatch.setenv("LQT_FIRST_CHUNK", "25")
    e0 = engine.Engine(tiny_dir, device=0)
    a0, c0 = e0.synthesize_tokens(ids, "en", max_new_tokens=40, seed=5, utterance_id=1)
    assert e0.stats().cooperative_launch == 1 and e0.stats().first_audio_ms < e0.stats().last_total_ms
    e0.close()
    monkeypatch.setenv("LQT_FK_NOCOOP", "1")
    e1 = engine.Engine(tiny_dir, device=0)
    a1, c1 = e1.synthesize_tokens(ids, "en", max_new_tokens=40, seed=5, utterance_id=1)
    s1 = e1.stats()
    assert s1.cooperative_launch == 0 and abs(s1.first_audio_ms - s1.last_total_ms) < 1e-3
    assert np.array_equal(c0, c1) and np.array_equal(a0, a1)
    e1.close()


@pytest.mark.parametrize("which,T,chunks", [("tiny", 40, [1, 2, 3, 5, 8, 13, 8]), ("tiny", 30, [30]), ("full", 110, [25, 25, 25, 25, 10]),
                                              ("full", 90, [7, 1, 40, 2, 40])])
def test_streaming_vocoder_is_bit_identical(request, which, T, chunks):
    """the vocoder decodes an utterance in chunks with carried state (conv halos, the 71-position K/V window of the
    pre-transformer, dilated-conv tails up to 54 rows) -- the concatenation must equal the one-shot decode BIT FOR BIT, for any
    chunking (single frames, chunks shorter than a layer's left context, chunks longer than the attention window)."""
    eng, m = pair(request, which)
    codes = np.random.default_rng(7 * T).integers(0, 2048, size=(T, 16))
    one_shot = eng.vocoder_decode(codes)
    streamed = eng.vocoder_stream(codes, chunks)
    assert streamed.shape == one_shot.shape
    assert np.array_equal(streamed, one_shot), (int(np.argmax(streamed != one_shot)), float(np.abs(streamed - one_shot).max()))
    # and a second utterance right after (state reset)
    codes2 = np.random.default_rng(T).integers(0, 2048, size=(12, 16))
    assert np.array_equal(eng.vocoder_stream(codes2, [5, 7]), eng.vocoder_decode(codes2))


def test_synthesize_stream_delivers_chunks_while_generating(request):
    """lqt_synthesize_stream on the full model: PCM arrives in chunks (4 frames first, then 2 s each), in order, covering the utterance exactly, equal to the
    returned buffer and to the non-streaming call; the first chunk arrives long before the call returns (generation of the
    remaining frames and vocoding of the finished ones overlap)."""
    import time
    eng, _ = pair(request, "full")
    orc = request.getfixturevalue("oracle_mod")
    ids = orc.wrap_text_ids(orc.synthetic_text_ids(20, 5))
    frames = 160
    got = []
    t0 = time.perf_counter()
    audio, codes = eng.synthesize_stream(ids, "en", 0.8, 50, 0.95, frames, 1234, 3, on_chunk=lambda first, pcm, t: got.append((first, pcm, t)))
    total = time.perf_counter() - t0
    assert codes.shape == (frames, 16) and audio.shape[0] == frames * 1920
    assert [g[0] for g in got] == sorted(g[0] for g in got) and got[0][0] == 0
    assert sum(g[1].shape[0] for g in got) == audio.shape[0] and [g[1].shape[0] // 1920 for g in got] == [4] + [25] * 6 + [6]
    assert np.array_equal(np.concatenate([g[1] for g in got]), audio)
    assert got[0][2] < 0.25 * total, (got[0][2], total)                 # first audio long before the end
    ref_audio, ref_codes = eng.synthesize_tokens(ids, "en", 0.8, 50, 0.95, frames, 1234, 3)
    assert np.array_equal(codes, ref_codes) and np.array_equal(audio, ref_audio)


def test_frame_kernel_resumes_across_launches(tiny_dir, monkeypatch):
    """the persistent kernel can stop after n frames and resume in a second launch (GenState + plain logits / last_hidden
    copies): $LQT_FK_SPLIT splits the launch; codes and logits must equal the single-launch run bit for bit"""
    from leaxer_qwen3_tts_b200 import engine
    ids = engine.wrap_text_ids([14990, 14615, 88225])
    e0 = engine.Engine(tiny_dir, device=0)
    prompt, trailing, pad = e0.build_prompt(ids, "en")
    c0, t0 = e0.generate(prompt, trailing, pad, e0.sampling(0.8, 50, 0.95, 14, 5, 1), trace=True)
    e0.close()
    monkeypatch.setenv("LQT_FK_SPLIT", "5")
    import subprocess, sys, os, json
    # the split point is read once per process: run the split variant in a child process
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from __graft_entry__ import load_package; load_package();"
            "from leaxer_qwen3_tts_b200 import engine; e = engine.Engine(%r, device=0); ids = engine.wrap_text_ids([14990, 14615, 88225]);"
            "p, t, d = e.build_prompt(ids, 'en'); c, tb = e.generate(p, t, d, e.sampling(0.8, 50, 0.95, 14, 5, 1), trace=True);"
            "np.savez(%r, c=c, tb=tb); e.close()") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), tiny_dir, "/tmp/lqt_split.npz")
    subprocess.run([sys.executable, "-c", code], check=True, env=dict(os.environ, LQT_FK_SPLIT="5"))
    g = np.load("/tmp/lqt_split.npz")
    assert np.array_equal(g["c"], c0) and np.array_equal(g["tb"], t0)
