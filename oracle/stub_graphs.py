"""TEST INFRASTRUCTURE -- numpy mirror of oracle/ort_shim/stub_graphs.h.

`StubModel` offers the graph interface that the host restatement in qwen3_tts_oracle.py drives
(text_project, codec_embed, code_predictor_embed, talker_prefill, talker_decode, code_predictor,
vocoder, speaker_encoder) but, like the ORT shim, computes deterministic hash outputs of the
contract's shapes and appends one trace line per call with the names, shapes and digests of the
tensors the REFERENCE would hand to Ort::Session::Run (src/tts_onnx.cpp:545-776). Running
oracle.synthesize_tokens(StubModel(), ..., schedule="reference") must therefore reproduce, line
for line, the trace of the reference's own compiled host (oracle/_ref/tts_host_ref):
tests/test_ref_host_pin.py. Nothing here is imported by the product.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

U32 = np.uint32


def mix32(x):
    """lowbias32 on a uint32 array (copy)"""
    x = np.array(x, dtype=U32, copy=True)
    x ^= x >> U32(16)
    x *= U32(0x7FEB352D)
    x ^= x >> U32(15)
    x *= U32(0x846CA68B)
    x ^= x >> U32(16)
    return x


def mix32s(x: int) -> int:
    return int(mix32(np.array([x & 0xFFFFFFFF], dtype=U32))[0])


def u2f(h):
    return (np.asarray(h, dtype=U32) >> U32(8)).astype(np.float32) * np.float32(2.0 ** -23) - np.float32(1.0)


def digest_words(w: np.ndarray, salt: int = 0):
    w = np.ascontiguousarray(w).view(U32).ravel()
    with np.errstate(over="ignore"):
        idx = np.arange(w.size, dtype=U32) * U32(0x9E3779B9) + U32(salt & 0xFFFFFFFF)
        h = mix32(w ^ mix32(idx))
        s = int(h.sum(dtype=np.uint64) & np.uint64(0xFFFFFFFF)) if h.size else 0
        x = int(np.bitwise_xor.reduce(h)) if h.size else 0
    return s, x


def key_of(d) -> int:
    s, x = d
    return mix32s(s ^ mix32s((x + 0x85EBCA6B) & 0xFFFFFFFF))


def fill_hash(n: int, base: int, scale: float) -> np.ndarray:
    with np.errstate(over="ignore"):
        j = np.arange(n, dtype=U32) + U32(base & 0xFFFFFFFF)
    return (np.float32(scale) * u2f(mix32(j))).astype(np.float32)


def _f32(t) -> np.ndarray:
    return np.ascontiguousarray(t.detach().numpy() if isinstance(t, torch.Tensor) else t, dtype=np.float32)


class StubModel:
    H, V, CPV, LAYERS, KVH, D, SPF = 1024, 3072, 2048, 28, 8, 128, 1920

    def __init__(self, eos_at: int = -1, has_speaker_encoder: bool = True):
        self.spec = SimpleNamespace(vocab=self.V, cp_vocab=self.CPV, cp_steps=15, hidden=self.H, layers=self.LAYERS,
                                    cp_layers=0, samples_per_frame=self.SPF)
        self.eos_at = eos_at
        self.has_speaker_encoder = has_speaker_encoder
        self.trace = []
        self.graph_calls = 0

    # ------------------------------------------------------------------------------------------
    def _log(self, graph, ins, outs):
        parts = [graph]
        for name, arr in ins:
            s, x = digest_words(arr)
            parts.append(f"{name}:{'x'.join(str(d) for d in arr.shape)}:{s:08x}:{x:08x}")
        self.trace.append(" ".join(parts) + " -> " + " ".join(outs))
        self.graph_calls += 1

    def _embed(self, graph, ids, salt):
        ids = np.asarray(ids, dtype=np.int64).reshape(1, -1)
        self._log(graph, [("input_ids", ids)], ["embeds"])
        rows = [fill_hash(self.H, mix32s((int(i) & 0xFFFFFFFF) ^ salt), 1.0) for i in ids[0]]
        return torch.from_numpy(np.stack(rows, 0))

    def text_project(self, ids):
        return self._embed("text_project", ids, 0x1111)

    def codec_embed(self, ids):
        return self._embed("codec_embed", ids, 0x2222)

    def code_predictor_embed(self, token: int, step: int):
        self._log("code_predictor_embed", [("input_ids", np.asarray([[token]], np.int64)),
                                           ("generation_step", np.asarray([step], np.int64))], ["embeds"])
        base = mix32s(mix32s((int(token) & 0xFFFFFFFF) ^ 0x3333) ^ ((int(step) * 0x9E3779B9) & 0xFFFFFFFF))
        return torch.from_numpy(fill_hash(self.H, base, 1.0))

    # ------------------------------------------------------------------------------------------
    def new_kv(self):
        return {"k": [torch.zeros(self.KVH, 0, self.D) for _ in range(self.LAYERS)],
                "v": [torch.zeros(self.KVH, 0, self.D) for _ in range(self.LAYERS)], "len": 0}

    def new_cp_kv(self):
        return {"len": 0}

    def _kv_names(self, prefix):
        out = []
        for i in range(self.LAYERS):
            out += [f"{prefix}_key_{i}", f"{prefix}_value_{i}"]
        return out

    def _talker(self, graph, rows, kv):
        rows = _f32(rows).reshape(-1, self.H)
        n_new, past = rows.shape[0], kv["len"]
        T = past + n_new
        mask = np.ones((1, T), np.int64)                       # the host's attention mask: all ones (:791, :843)
        ins = [("inputs_embeds", rows.reshape(1, n_new, self.H)), ("attention_mask", mask)]
        salt = 0x4444
        if graph == "talker_decode":
            pk = 0x9999
            names = self._kv_names("past")
            for i in range(self.LAYERS):
                for j, kind in enumerate(("k", "v")):
                    arr = _f32(kv[kind][i]).reshape(1, self.KVH, past, self.D)
                    ins.append((names[2 * i + j], arr))
                    pk = mix32s(pk ^ key_of(digest_words(arr, 2 + 2 * i + j)))
            salt = pk
        self._log(graph, ins, ["logits", "last_hidden"] + self._kv_names("present"))
        ck, prev = [], salt
        for p in range(n_new):
            prev = mix32s(prev ^ key_of(digest_words(rows[p], past + p)))
            ck.append(prev)
        logits = np.stack([fill_hash(self.V, c, 4.0) for c in ck], 0)
        if graph == "talker_decode" and T == self.eos_at:
            logits[0, 2150] = np.float32(100.0)
        hidden = fill_hash(self.H, ck[-1] ^ 0x55555555, 1.0)
        for i in range(2 * self.LAYERS):
            new = np.empty((self.KVH, n_new, self.D), np.float32)
            for h in range(self.KVH):
                for p in range(n_new):
                    new[h, p] = fill_hash(self.D, (ck[p] + 0x10000 * i + 128 * h + 0x777) & 0xFFFFFFFF, 1.0)
            kind = "k" if i % 2 == 0 else "v"
            kv[kind][i // 2] = torch.cat([kv[kind][i // 2], torch.from_numpy(new)], dim=1)
        kv["len"] = T
        return torch.from_numpy(logits), torch.from_numpy(hidden)

    def talker_prefill(self, embeds, kv):
        return self._talker("talker_prefill", embeds, kv)

    def talker_decode(self, embed, kv):
        logits, hidden = self._talker("talker_decode", embed, kv)
        return logits[0], hidden

    def code_predictor(self, embeds, step: int):
        rows = _f32(embeds).reshape(-1, self.H)
        self._log("code_predictor", [("inputs_embeds", rows.reshape(1, -1, self.H)),
                                     ("generation_step", np.asarray([step], np.int64))], ["logits"])
        prev = 0x6666
        for p in range(rows.shape[0]):
            prev = mix32s(prev ^ key_of(digest_words(rows[p], p)))
        return torch.from_numpy(fill_hash(self.CPV, prev ^ ((int(step) * 0x9E3779B9) & 0xFFFFFFFF), 4.0))

    def vocoder(self, codes):
        codes = np.ascontiguousarray(np.asarray(codes, dtype=np.int64).reshape(1, -1, 16))
        self._log("tokenizer12hz_decode", [("audio_codes", codes)], ["audio_values", "lengths"])
        n = codes.shape[1] * self.SPF
        return torch.from_numpy(fill_hash(n, key_of(digest_words(codes)) ^ 0x8888, 0.5)), n

    def speaker_encoder(self, mel_t):
        mel_t = _f32(mel_t)
        arr = mel_t.reshape(1, mel_t.shape[0], mel_t.shape[1])
        self._log("speaker_encoder", [("mel", arr)], ["embedding"])
        return torch.from_numpy(fill_hash(self.H, key_of(digest_words(arr)) ^ 0xAAAA, 1.0))
