"""CPU ORACLE for the Qwen3-TTS hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file. The product path (csrc/ + engine.py + host/) never does and has no CPU fallback.

What it restates (all file:line relative to /root/reference/):
  * host orchestration, line for line: id wrapping src/tts_onnx.cpp:238-262, prompt assembly
    :442-539, frame loop A :782-849, sub-code loop B :851-872, sampler :878-950, vocoder length
    handling :759-776;
  * the seven (+1) graphs behind Ort::Session::Run, by their I/O contract (:545-776) with the
    internal architecture frozen in leaxer-qwen3-tts_b200/modelspec.py (the .onnx graphs are a
    third-party, un-vendored artefact: HF zukky/Qwen3-TTS-ONNX-DLL `onnx/onnx_kv_06b`, no pinned
    revision, README.md:69-80; run by ONNX Runtime 1.20.0, .github/workflows/ci.yml:10).

PARITY PINNING STATUS
  * graph arithmetic: **parity unpinned** -- the reference ships no golden logits/tokens/audio
    (tests/test_onnx.cpp only checks constants), ONNX Runtime and the graphs are absent here, so
    nothing external can pin the numbers. Architecture follows the upstream Qwen3-TTS /
    Qwen3-Omni-talker family (transformers/models/qwen3_omni_moe/modeling_qwen3_omni_moe.py
    :2309-2481, :3283-3366, :3645-3778).
  * host orchestration + sampler filters: PINNED against the reference's own src/tts_onnx.cpp,
    compiled unmodified against a stand-in onnxruntime_cxx_api.h (oracle/ort_shim/) whose
    Session::Run dispatches to deterministic stub graphs: oracle/Makefile builds
    oracle/_ref/tts_host_ref, tests/test_ref_host_pin.py runs this file's host logic over the same
    stubs (oracle/stub_graphs.py) and requires the identical per-call trace (graph order, tensor
    names, shapes, input digests: prompt rows, masks, KV round trip, trailing schedule, flattened
    codes), identical tokens with --top-k 1, and bit-identical top-k / softmax / top-p results.

Numerics: fp32 activations, weights are the bf16 values stored in the .lqw files upcast to fp32,
talker K (post-norm, post-RoPE) and V rounded to bf16 when they enter the KV cache (north_star:
"paged bf16 KV cache"); the code predictor's 17-position KV stays fp32. Sampler: reference
semantics with the RNG replaced by Philox4x32-10 (north_star), exp evaluated
in float64 and rounded to float32 so that CPU and GPU agree bit for bit.
"""
from __future__ import annotations

import math
import os
import sys
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from __graft_entry__ import load_package  # noqa: E402

_pkg = load_package()
from leaxer_qwen3_tts_b200 import modelspec as ms  # noqa: E402

# ---- reference constants (src/tts_onnx.h:39-69) ------------------------------------------------
TTS_BOS, TTS_EOS, TTS_PAD = 151672, 151673, 151671
IM_START, IM_END, ASSISTANT = 151644, 151645, 77091
CODEC_BOS, CODEC_EOS, CODEC_PAD = 2149, 2150, 2148
CODEC_THINK, CODEC_NOTHINK, CODEC_THINK_BOS, CODEC_THINK_EOS = 2154, 2155, 2156, 2157
LANG_IDS = {"auto": 0, "en": 2050, "zh": 2051, "ja": 2052, "ko": 2053}   # tts_onnx.h:59-62, 230-238
SAMPLE_RATE = 24000


@dataclass
class SamplingParams:                      # src/tts_onnx.h:99-105
    temperature: float = 0.8
    top_p: float = 0.95
    top_k: int = 50
    repetition_penalty: float = 1.0        # declared, never read (tts_onnx.h:103)
    max_new_tokens: int = 2048
    # engine extension (north_star: seeded Philox sampler); greedy = argmax, lowest index on ties
    seed: int = 0
    utterance_id: int = 0
    greedy: bool = False


# ================================================================================================
# Philox4x32-10 and the sampler (src/tts_onnx.cpp:878-950)
# ================================================================================================
_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    c0, c1, c2, c3 = [int(x) & _MASK for x in counter]
    k0, k1 = [int(x) & _MASK for x in key]
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        hi0, lo0 = p0 >> 32, p0 & _MASK
        hi1, lo1 = p1 >> 32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0, k1 = (k0 + _W0) & _MASK, (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def philox_uniform(seed: int, utterance_id: int, frame: int, codebook: int) -> np.float32:
    """u in [0,1): key = (seed, utterance id), counter = (frame, codebook, 0, 0), 24 bits of x0."""
    x0 = philox4x32_10((frame, codebook, 0, 0), (seed, utterance_id))[0]
    return np.float32(x0 >> 8) * np.float32(2.0 ** -24)


def _exp_f32_via_f64(x: np.ndarray) -> np.ndarray:
    return np.exp(x.astype(np.float64)).astype(np.float32)


def sampler_filtered_probs(logits: np.ndarray, p: SamplingParams):
    """Steps 2-5 of Appendix C. Returns float32 probs[V] after temperature / top-k / softmax /
    top-p / renormalise, exactly as src/tts_onnx.cpp:878-898 computes them (f32, serial sums)."""
    x = np.array(logits, dtype=np.float32, copy=True)
    V = x.shape[0]
    if p.temperature > 0.0 and np.float32(p.temperature) != np.float32(1.0):      # :882-884
        x = x / np.float32(p.temperature)
    if 0 < p.top_k < V:                                                          # :917-927
        thr = np.partition(x, V - p.top_k)[V - p.top_k]      # k-th largest value
        x = np.where(x < thr, np.float32(-np.inf), x)
    m = x.max()                                                                  # :907-915
    e = np.where(np.isneginf(x), np.float32(0), _exp_f32_via_f64(x - m)).astype(np.float32)
    s = np.float32(0)
    for v in e[e > 0]:                       # left-to-right f32 sum (zeros do not change it)
        s = np.float32(s + v)
    prob = (e / s).astype(np.float32)
    if p.top_p < 1.0:                                                            # :929-950, :893-898
        nz = np.nonzero(prob > 0)[0]
        order = nz[np.argsort(-prob[nz], kind="stable")]      # prob desc, index asc on ties
        c = np.float32(0)
        cutoff = len(order)
        for i, idx in enumerate(order):
            c = np.float32(c + prob[idx])
            if c > np.float32(p.top_p):
                cutoff = i + 1
                break
        prob[order[cutoff:]] = 0
        s2 = np.float32(0)
        for v in prob[prob > 0]:
            s2 = np.float32(s2 + v)
        if s2 > 0:
            prob = (prob / s2).astype(np.float32)
    return prob


def sample_token(logits: np.ndarray, p: SamplingParams, frame: int, codebook: int) -> int:
    """src/tts_onnx.cpp:878-905 with the draw re-defined on Philox: smallest i with cdf[i] > u."""
    if p.greedy:
        return int(np.argmax(logits))        # first maximum = lowest index
    prob = sampler_filtered_probs(logits, p)
    u = philox_uniform(p.seed, p.utterance_id, frame, codebook)
    c = np.float32(0)
    last = 0
    for i in np.nonzero(prob > 0)[0]:
        c = np.float32(c + prob[i])
        last = int(i)
        if c > u:
            return int(i)
    return last


# ================================================================================================
# weights
# ================================================================================================
def _t(a: np.ndarray) -> torch.Tensor:
    if a.dtype == np.uint16:
        return torch.from_numpy(ms.bf16_bits_to_f32(np.asarray(a)))
    return torch.from_numpy(np.array(a, dtype=np.float32))


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


class _Graph:
    """Lazy fp32 view of one .lqw file's tensors."""

    def __init__(self, tensors):
        self._raw, self._cache = tensors, {}

    def __getitem__(self, k) -> torch.Tensor:
        if k not in self._cache:
            self._cache[k] = _t(self._raw[k])
        return self._cache[k]

    def __contains__(self, k):
        return k in self._raw

    def rows(self, k, ids) -> torch.Tensor:
        """gather rows of a (possibly huge) bf16 table without converting all of it"""
        return _t(np.asarray(self._raw[k])[np.asarray(ids, dtype=np.int64)])


def rmsnorm(x, w, eps):
    var = x.pow(2).mean(-1, keepdim=True)
    return x * torch.rsqrt(var + eps) * w


def rope_apply(x, cos, sin):
    """x [n, heads, D]; cos/sin [n, D/2]; rotate-half convention."""
    h = x.shape[-1] // 2
    x1, x2 = x[..., :h], x[..., h:]
    c, s = cos[:, None, :], sin[:, None, :]
    return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1)


class OracleModel:
    def __init__(self, model_dir: str, kv_bf16: bool = True):
        self.model_dir = model_dir
        self.kv_bf16 = kv_bf16          # talker KV rounding point (False mirrors LQT_KV_F32 parity mode)
        self.spec, graphs = ms.load_model_dir(model_dir)
        missing = [g for g in ms.GRAPH_FILES if g not in graphs]
        if missing:
            raise FileNotFoundError(f"missing graph files: {missing}")     # tts_onnx.cpp:100-104
        self.g = {k: _Graph(v) for k, v in graphs.items()}
        self.has_speaker_encoder = "speaker_encoder" in graphs
        self.graph_calls = 0

    # ---------------------------------------------------------------- embeddings (:545-613)
    def text_project(self, ids) -> torch.Tensor:
        """input_ids i64 [1,S] -> embeds f32 [S,H]"""
        self.graph_calls += 1
        g = self.g["text_project"]
        e = g.rows("embed", ids)
        h = F.silu(e @ g["fc1.weight"].T + g["fc1.bias"])
        return h @ g["fc2.weight"].T + g["fc2.bias"]

    def codec_embed(self, ids) -> torch.Tensor:
        self.graph_calls += 1
        return self.g["codec_embed"].rows("embed", ids)

    def code_predictor_embed(self, token: int, step: int) -> torch.Tensor:
        self.graph_calls += 1
        return self.g["code_predictor_embed"]["embed"][step, token].clone()

    # ---------------------------------------------------------------- transformer core
    def _layer(self, g, pre, x, pos0, K, V, eps, n_heads, n_kv, D, cos, sin, kv_bf16,
               window=None, ls=False, qk_norm=True):
        n = x.shape[0]
        qd, kvd = n_heads * D, n_kv * D
        h = rmsnorm(x, g[f"{pre}.ln1"], eps)
        qkv = h @ g[f"{pre}.wqkv"].T
        q = qkv[:, :qd].reshape(n, n_heads, D)
        k = qkv[:, qd:qd + kvd].reshape(n, n_kv, D)
        v = qkv[:, qd + kvd:].reshape(n, n_kv, D)
        if qk_norm:
            q = rmsnorm(q, g[f"{pre}.qnorm"], eps)
            k = rmsnorm(k, g[f"{pre}.knorm"], eps)
        c, s = cos[pos0:pos0 + n], sin[pos0:pos0 + n]
        q, k = rope_apply(q, c, s), rope_apply(k, c, s)
        if kv_bf16:
            k, v = bf16_round(k), bf16_round(v)
        K = k if K is None else torch.cat([K, k], 0)          # [t, n_kv, D]
        V = v if V is None else torch.cat([V, v], 0)
        t = K.shape[0]
        rep = n_heads // n_kv
        Kh = K.repeat_interleave(rep, dim=1)                  # [t, n_heads, D]
        Vh = V.repeat_interleave(rep, dim=1)
        scores = torch.einsum("nhd,thd->hnt", q, Kh) * (D ** -0.5)
        qi = torch.arange(pos0, pos0 + n)[:, None]
        kj = torch.arange(t)[None, :]
        mask = kj <= qi
        if window is not None:
            mask = mask & (qi - kj < window)
        scores = scores.masked_fill(~mask[None], float("-inf"))
        p = torch.softmax(scores, dim=-1)
        o = torch.einsum("hnt,thd->nhd", p, Vh).reshape(n, qd)
        a = o @ g[f"{pre}.wo"].T
        x = x + (g[f"{pre}.ls1"] * a if ls else a)
        h2 = rmsnorm(x, g[f"{pre}.ln2"], eps)
        m = (F.silu(h2 @ g[f"{pre}.wgate"].T) * (h2 @ g[f"{pre}.wup"].T)) @ g[f"{pre}.wdown"].T
        x = x + (g[f"{pre}.ls2"] * m if ls else m)
        return x, K, V

    def _talker(self, x, kv):
        """x [n,H] new rows appended after kv['len'] positions. Returns logits [n,V], hidden [n,H]."""
        sp, g = self.spec, self.g["talker_prefill"]
        pos0 = kv["len"]
        for i in range(sp.layers):
            x, kv["k"][i], kv["v"][i] = self._layer(
                g, f"l{i}", x, pos0, kv["k"][i], kv["v"][i], sp.rms_eps, sp.heads, sp.kv_heads,
                sp.head_dim, g["rope_cos"], g["rope_sin"], kv_bf16=self.kv_bf16)
        kv["len"] = pos0 + x.shape[0]
        hid = rmsnorm(x, g["norm"], sp.rms_eps)
        return hid @ g["head"].T, hid

    def new_kv(self):
        return {"k": [None] * self.spec.layers, "v": [None] * self.spec.layers, "len": 0}

    # talker_prefill.onnx (:615-665): logits [P,V], last_hidden = final-norm hidden of last position
    def talker_prefill(self, embeds: torch.Tensor, kv):
        self.graph_calls += 1
        logits, hid = self._talker(embeds, kv)
        return logits, hid[-1]

    # talker_decode.onnx (:667-732)
    def talker_decode(self, embed: torch.Tensor, kv):
        self.graph_calls += 1
        logits, hid = self._talker(embed[None, :], kv)
        return logits[0], hid[0]

    # code_predictor.onnx (:734-757): full re-forward over L rows, head `step` on the last row
    def code_predictor(self, embeds: torch.Tensor, step: int) -> torch.Tensor:
        self.graph_calls += 1
        kv = self.new_cp_kv()
        hid = self._cp_rows(embeds, kv)
        return hid[-1] @ self.g["code_predictor"]["heads"][step].T

    def new_cp_kv(self):
        return {"k": [None] * self.spec.cp_layers, "v": [None] * self.spec.cp_layers, "len": 0}

    def _cp_rows(self, x, kv):
        """KV-cached form of the same graph (identical arithmetic; predictor KV stays fp32)."""
        sp, g = self.spec, self.g["code_predictor"]
        if "in_proj.weight" in g:
            x = x @ g["in_proj.weight"].T + g["in_proj.bias"]
        pos0 = kv["len"]
        for i in range(sp.cp_layers):
            x, kv["k"][i], kv["v"][i] = self._layer(
                g, f"l{i}", x, pos0, kv["k"][i], kv["v"][i], sp.rms_eps, sp.cp_heads,
                sp.cp_kv_heads, sp.head_dim, g["rope_cos"], g["rope_sin"], kv_bf16=False)
        kv["len"] = pos0 + x.shape[0]
        return rmsnorm(x, g["norm"], sp.rms_eps)

    def code_predictor_cached(self, new_rows: torch.Tensor, step: int, kv) -> torch.Tensor:
        hid = self._cp_rows(new_rows, kv)
        return hid[-1] @ self.g["code_predictor"]["heads"][step].T

    # ---------------------------------------------------------------- vocoder (:759-776)
    @staticmethod
    def _causal_conv(x, w, b, dilation=1):
        """x [L,Cin] channels-last; w [Cout,taps,Cin]; left pad (taps-1)*dilation, no right pad."""
        taps = w.shape[1]
        xt = F.pad(x.T[None], ((taps - 1) * dilation, 0))
        return F.conv1d(xt, w.permute(0, 2, 1).contiguous(), b, dilation=dilation)[0].T

    @staticmethod
    def _tconv(x, w, b):
        """x [L,Cin]; w [s,Cout,nh,Cin] (nh=1: kernel=s ; nh=2: kernel=2s); output [L*s,Cout]:
        full transposed conv trimmed on the RIGHT by (kernel - stride) -> exactly L*s, causal."""
        s, cout, nh, cin = w.shape
        wt = w.permute(3, 1, 2, 0).reshape(cin, cout, nh * s)          # [Cin,Cout,k], k = h*s + r
        y = F.conv_transpose1d(x.T[None], wt.contiguous(), b, stride=s)[0]
        return y[:, : x.shape[0] * s].T

    @staticmethod
    def _snake(x, alpha, beta):
        return x + (1.0 / (torch.exp(beta) + 1e-9)) * torch.sin(x * torch.exp(alpha)).pow(2)

    def vocoder_stages(self, codes):
        """audio_codes i64 [T,16] -> dict of intermediate activations (channels-last)."""
        sp, g = self.spec, self.g["tokenizer12hz_decode"]
        codes = torch.as_tensor(np.asarray(codes, dtype=np.int64)).reshape(-1, sp.cp_steps + 1)
        T = codes.shape[0]
        st = {}
        sem = g["rvq.sem.codebook"][0][codes[:, 0]]
        aco = torch.zeros_like(sem)
        for j in range(sp.cp_steps):                       # gather-SUM in codebook order
            aco = aco + g["rvq.aco.codebook"][j][codes[:, j + 1]]
        x = sem @ g["rvq.sem.out_proj"].T + aco @ g["rvq.aco.out_proj"].T
        st["rvq"] = x
        x = self._causal_conv(x, g["pre_conv.weight"], g["pre_conv.bias"])
        st["pre_conv"] = x
        for i in range(sp.voc_layers):
            x, _, _ = self._layer(g, f"pt.l{i}", x, 0, None, None, sp.voc_rms_eps, sp.voc_heads,
                                  sp.voc_heads, sp.voc_head_dim, g["pt.rope_cos"], g["pt.rope_sin"],
                                  kv_bf16=False, window=sp.voc_window, ls=True, qk_norm=False)
        x = rmsnorm(x, g["pt.norm"], sp.voc_rms_eps)
        st["pre_transformer"] = x
        for u in range(len(sp.voc_upsampling_ratios)):
            x = self._tconv(x, g[f"up{u}.tconv.weight"], g[f"up{u}.tconv.bias"])
            r = x
            h = F.conv1d(F.pad(x.T[None], (6, 0)), g[f"up{u}.dw.weight"].T[:, None, :].contiguous(),
                         g[f"up{u}.dw.bias"], groups=x.shape[1])[0].T
            h = F.layer_norm(h, (h.shape[1],), g[f"up{u}.ln.weight"], g[f"up{u}.ln.bias"], 1e-6)
            h = F.gelu(h @ g[f"up{u}.pw1.weight"].T + g[f"up{u}.pw1.bias"])
            h = h @ g[f"up{u}.pw2.weight"].T + g[f"up{u}.pw2.bias"]
            x = r + g[f"up{u}.gamma"] * h
            st[f"up{u}"] = x
        x = self._causal_conv(x, g["dec.conv_in.weight"], g["dec.conv_in.bias"])
        st["dec.conv_in"] = x
        for b in range(len(sp.voc_upsample_rates)):
            x = self._snake(x, g[f"dec.b{b}.snake.alpha"], g[f"dec.b{b}.snake.beta"])
            x = self._tconv(x, g[f"dec.b{b}.tconv.weight"], g[f"dec.b{b}.tconv.bias"])
            for r, dil in enumerate((1, 3, 9)):
                p = f"dec.b{b}.r{r}"
                h = self._snake(x, g[f"{p}.snake1.alpha"], g[f"{p}.snake1.beta"])
                h = self._causal_conv(h, g[f"{p}.conv1.weight"], g[f"{p}.conv1.bias"], dil)
                h = self._snake(h, g[f"{p}.snake2.alpha"], g[f"{p}.snake2.beta"])
                h = self._causal_conv(h, g[f"{p}.conv2.weight"], g[f"{p}.conv2.bias"])
                x = x + h
            st[f"dec.b{b}"] = x
        x = self._snake(x, g["dec.snake_out.alpha"], g["dec.snake_out.beta"])
        x = self._causal_conv(x, g["dec.conv_out.weight"], g["dec.conv_out.bias"])
        st["audio"] = x[:, 0].clamp(-1.0, 1.0)
        assert st["audio"].shape[0] == T * sp.samples_per_frame
        return st

    def vocoder(self, codes):
        """-> (audio_values f32 [T*1920], lengths) ; the host reads lengths[0] samples (:771-775)"""
        self.graph_calls += 1
        a = self.vocoder_stages(codes)["audio"]
        return a, a.shape[0]

    # ---------------------------------------------------------------- speaker encoder (:367-403)
    def speaker_encoder(self, mel_t: torch.Tensor) -> torch.Tensor:
        """mel_t f32 [frames,128] (already transposed by the host) -> [H]"""
        self.graph_calls += 1
        g = self.g["speaker_encoder"]

        def conv_same(x, w, b):
            k = w.shape[1]
            return F.conv1d(x.T[None], w.permute(0, 2, 1).contiguous(), b, padding=k // 2)[0].T
        x = F.relu(conv_same(mel_t, g["in_conv.weight"], g["in_conv.bias"]))
        for i in range(self.spec.spk_layers):
            x = x + F.relu(conv_same(x, g[f"l{i}.conv.weight"], g[f"l{i}.conv.bias"]))
        mean = x.mean(0)
        std = (x - mean).pow(2).mean(0).add(1e-5).sqrt()
        return torch.cat([mean, std]) @ g["fc.weight"].T + g["fc.bias"]


# ================================================================================================
# Host orchestration restated (src/tts_onnx.cpp:238-539, 782-872)
# ================================================================================================
def wrap_text_ids(text_ids):
    """src/tts_onnx.cpp:243-259"""
    return [IM_START, ASSISTANT, TTS_BOS] + [int(t) for t in text_ids] + [TTS_EOS, IM_END]


@dataclass
class UtteranceState:
    kv: dict
    last_hidden: torch.Tensor = None
    trailing_text_hidden: torch.Tensor = None     # [trailing_len, H]
    tts_pad_embed: torch.Tensor = None
    trailing_len: int = 0
    trace: dict = field(default_factory=dict)


def build_prompt_embeddings(m: OracleModel, input_ids, lang: str, st: UtteranceState,
                            speaker_embed=None) -> torch.Tensor:
    """src/tts_onnx.cpp:442-539. Returns prompt [P,H]; fills st.trailing_*, st.tts_pad_embed."""
    has_spk = speaker_embed is not None and len(speaker_embed) > 0
    tts = m.text_project([TTS_BOS, TTS_EOS, TTS_PAD])                        # :459-463
    tts_bos, tts_eos, st.tts_pad_embed = tts[0], tts[1], tts[2]
    if lang == "auto":                                                        # :466-476
        codec_prefill = [CODEC_NOTHINK, CODEC_THINK_BOS, CODEC_THINK_EOS]
    else:
        codec_prefill = [CODEC_THINK, CODEC_THINK_BOS, LANG_IDS[lang], CODEC_THINK_EOS]
    codec_prefill += [CODEC_PAD, CODEC_BOS]
    codec_embeds = m.codec_embed(codec_prefill)                               # :478
    if has_spk:                                                               # :481-490
        se = torch.as_tensor(np.asarray(speaker_embed, dtype=np.float32))[None, :]
        codec_embeds = torch.cat([codec_embeds[:-1], se, codec_embeds[-1:]], 0)
    role = m.text_project(list(input_ids[:3]))                                # :493-494
    pad_count = len(codec_prefill) - 2 + (1 if has_spk else 0)                # :497-498
    text_part = torch.cat([st.tts_pad_embed[None].repeat(pad_count, 1), tts_bos[None]], 0)
    talker = text_part + codec_embeds[:pad_count + 1]                         # :511-512
    text_start, text_end = 3, len(input_ids) - 2                              # :515-516
    first = m.text_project([input_ids[text_start]])[0] + codec_embeds[pad_count + 1]   # :518-520
    prompt = torch.cat([role, talker, first[None]], 0)                        # :523-527
    rows = [m.text_project([input_ids[i]])[0] for i in range(text_start + 1, text_end)]  # :531-534
    rows.append(tts_eos)
    st.trailing_text_hidden = torch.stack(rows, 0)
    st.trailing_len = st.trailing_text_hidden.shape[0]                        # :536
    return prompt


def predict_subcodes(m: OracleModel, code0: int, st: UtteranceState, p: SamplingParams, frame: int,
                     schedule: str = "cached", trace=None):
    """src/tts_onnx.cpp:851-872. schedule='reference' re-runs the graph on the growing sequence
    exactly like the reference; 'cached' is the arithmetic-identical KV-cached form."""
    first = m.codec_embed([code0])[0]                                         # :854
    seq = torch.stack([st.last_hidden, first], 0)                             # :857-860
    sub = []
    kv = m.new_cp_kv()
    new_rows = seq
    for j in range(m.spec.cp_steps):                                          # :862-869
        if schedule == "reference":
            logits = m.code_predictor(seq, j)
        else:
            logits = m.code_predictor_cached(new_rows, j, kv)
        if trace is not None:
            trace.append(logits.numpy().copy())
        tok = sample_token(logits.numpy(), p, frame, j + 1)
        sub.append(tok)
        e = m.code_predictor_embed(tok, j)
        seq = torch.cat([seq, e[None]], 0)
        new_rows = e[None]
    return sub


def generate_codes(m: OracleModel, prompt: torch.Tensor, st: UtteranceState, p: SamplingParams,
                   schedule: str = "cached", forced_codes=None, trace: dict | None = None):
    """src/tts_onnx.cpp:782-849. forced_codes ([T,16]) = teacher forcing for parity triage.
    trace (optional dict) collects per-frame logits / hidden for golden fixtures."""
    V = m.spec.vocab
    logits_all, st.last_hidden = m.talker_prefill(prompt, st.kv)              # :794
    last_logits = logits_all[-1].clone()                                      # :797-798
    all_codes = []
    if trace is not None:
        trace.update({"talker_logits": [], "last_hidden": [], "cp_logits": [], "next_in": []})
    for step in range(p.max_new_tokens):                                      # :801
        last_logits[2048:V] = torch.where(
            torch.arange(2048, V) == CODEC_EOS, last_logits[2048:V],
            torch.tensor(float("-inf")))                                      # :803-807
        if trace is not None:
            trace["talker_logits"].append(last_logits.numpy().copy())
            trace["last_hidden"].append(st.last_hidden.numpy().copy())
        code0 = sample_token(last_logits.numpy(), p, step, 0)                 # :810
        if forced_codes is not None:
            if step >= len(forced_codes):
                break
            code0 = int(forced_codes[step][0])
        if code0 == CODEC_EOS:                                                # :812
            break
        cp_trace = [] if trace is not None else None
        sub = predict_subcodes(m, code0, st, p, step, schedule, cp_trace)     # :815
        if forced_codes is not None:
            sub = [int(c) for c in forced_codes[step][1:]]
        if trace is not None:
            trace["cp_logits"].append(np.stack(cp_trace))
        all_codes.append([code0] + sub)                                       # :818-821
        e = m.codec_embed([code0])[0].clone()                                 # :824
        for i in range(m.spec.cp_steps):                                      # :825-830
            e = e + m.code_predictor_embed(sub[i], i)
        if step < st.trailing_len:                                            # :833-842
            e = e + st.trailing_text_hidden[step]
        else:
            e = e + st.tts_pad_embed
        if trace is not None:
            trace["next_in"].append(e.numpy().copy())
        if schedule == "reference":
            # the reference hands the whole KV cache to the graph and copies it back (:684-729)
            st.kv["k"] = [k.clone() for k in st.kv["k"]]
            st.kv["v"] = [v.clone() for v in st.kv["v"]]
        last_logits, st.last_hidden = m.talker_decode(e, st.kv)               # :845
        last_logits = last_logits.clone()
    return all_codes


def synthesize_tokens(m: OracleModel, token_ids, lang: str = "auto", p: SamplingParams = None,
                      speaker_embed=None, schedule: str = "cached", forced_codes=None,
                      trace: dict | None = None, run_vocoder: bool = True):
    """src/tts_onnx.cpp:405-436 (and :299-313 for the clone variant). -> (audio f32 np, codes)"""
    p = p or SamplingParams()
    st = UtteranceState(kv=m.new_kv())                                        # :411
    prompt = build_prompt_embeddings(m, token_ids, lang, st, speaker_embed)   # :414
    if trace is not None:
        trace["prompt"] = prompt.numpy().copy()
        trace["trailing"] = st.trailing_text_hidden.numpy().copy()
        trace["tts_pad"] = st.tts_pad_embed.numpy().copy()
    codes = generate_codes(m, prompt, st, p, schedule, forced_codes, trace)   # :417
    if not codes:                                                             # :418
        return np.zeros(0, np.float32), np.zeros((0, 16), np.int64)
    codes_np = np.asarray(codes, dtype=np.int64)                              # :421-427
    if not run_vocoder:
        return None, codes_np
    audio, n = m.vocoder(codes_np)                                            # :430
    return audio.numpy()[:n].copy(), codes_np


def extract_speaker_embedding(m, mel: np.ndarray) -> np.ndarray:
    """src/tts_onnx.cpp:367-403 (run_speaker_encoder): the log-mel of the reference clip arrives [num_mels, frames]
    row-major from io::MelExtractor (:331-359) and is transposed to [1, frames, num_mels] for the graph."""
    mel = np.asarray(mel, dtype=np.float32)
    assert mel.ndim == 2 and mel.shape[0] == 128                              # :371
    mel_t = np.ascontiguousarray(mel.T)                                       # :375-380
    return m.speaker_encoder(torch.from_numpy(mel_t)).numpy().copy()


def synthetic_text_ids(n_text: int, seed: int = 1234):
    """Synthetic prompt ids uniform in [0,151643); hash-based, machine independent."""
    u = ms.uniform_pm1(seed, "synthetic_text_ids", n_text)
    return [int(x) for x in ((u.astype(np.float64) + 1.0) * 0.5 * 151643).astype(np.int64)]
