"""ctypes binding over the C-ABI (include/lqt_b200.h). This is the call surface tests and bench.py
use; it mirrors leaxer_qwen::TTSEngine's private run_* graph runners and public synthesize_tokens
(/root/reference/src/tts_onnx.h:118-226). No CPU fallback: a missing library or GPU raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LQT_B200_LIB") or os.path.join(_HERE, "csrc", "liblqt_b200.so")   # $LQT_B200_LIB: the profiling build

LANG_CODEC_ID = {"auto": 0, "en": 2050, "zh": 2051, "ja": 2052, "ko": 2053}   # tts_onnx.h:230-238

# every symbol include/lqt_b200.h declares
SYMBOLS = [
    "lqt_create", "lqt_create_ex", "lqt_create_error", "lqt_destroy", "lqt_last_error", "lqt_get_info", "lqt_get_stats",
    "lqt_reset_stats", "lqt_text_project", "lqt_codec_embed", "lqt_code_predictor_embed",
    "lqt_talker_prefill", "lqt_talker_decode", "lqt_kv_reset", "lqt_kv_len", "lqt_code_predictor",
    "lqt_vocoder_decode", "lqt_speaker_encoder", "lqt_sample", "lqt_generate", "lqt_synthesize_tokens",
    "lqt_build_prompt", "lqt_debug_timeline", "lqt_debug_exchange", "lqt_check_model_file",
    "lqt_synthesize_batch", "lqt_debug_tc_gemm", "lqt_log_mel", "lqt_speaker_embed_audio",
    "lqt_vocoder_stream_reset", "lqt_vocoder_stream_chunk", "lqt_synthesize_stream",
]


class Sampling(C.Structure):
    _fields_ = [("temperature", C.c_float), ("top_p", C.c_float), ("top_k", C.c_int32),
                ("max_new_tokens", C.c_int32), ("seed", C.c_uint32), ("utterance_id", C.c_uint32),
                ("greedy", C.c_int32)]


class Options(C.Structure):
    _fields_ = [("kv_dtype", C.c_int32), ("n_slots", C.c_int32), ("frame_impl", C.c_int32)]


class Info(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "hidden", "layers", "heads", "kv_heads", "head_dim", "vocab", "cp_vocab", "cp_steps",
        "samples_per_frame", "sample_rate", "has_speaker_encoder", "max_pos", "num_sms")]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("graph_launches", C.c_uint64),
                ("last_generate_ms", C.c_float), ("last_vocoder_ms", C.c_float),
                ("last_prefill_ms", C.c_float), ("last_frames", C.c_int32),
                ("last_total_ms", C.c_float), ("first_audio_ms", C.c_float),
                ("frame_impl_active", C.c_int32), ("cooperative_launch", C.c_int32)]


FRAME_IMPL = {"persistent": 0, "graph": 1, "auto": 2, "batched": 3}      # include/lqt_b200.h LQT_FRAME_*


class BatchRequest(C.Structure):
    _fields_ = [("token_ids", C.c_void_p), ("n_ids", C.c_int32), ("lang_codec_id", C.c_int32),
                ("speaker_embed", C.c_void_p), ("utterance_id", C.c_uint32), ("max_new_tokens", C.c_int32),
                ("forced_codes", C.c_void_p), ("n_forced", C.c_int32),
                ("audio_out", C.c_void_p), ("audio_capacity", C.c_int64), ("n_samples", C.POINTER(C.c_int64)),
                ("codes_out", C.c_void_p), ("n_frames", C.POINTER(C.c_int32)), ("logits_trace", C.c_void_p)]


class BatchOptions(C.Structure):
    _fields_ = [("max_concurrent", C.c_int32), ("planes", C.c_int32), ("poll_frames", C.c_int32)]


AUDIO_CB = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_float), C.c_int64, C.c_int64)

_lib = None


def load_library():
    """Loads liblqt_b200.so and checks that it exports every declared symbol. Raises if not."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                           f"g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    missing = [s for s in SYMBOLS if not hasattr(lib, s)]
    if missing:
        raise RuntimeError(f"liblqt_b200.so lacks symbols: {missing}")
    lib.lqt_create.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.lqt_create_ex.argtypes = [C.c_char_p, C.c_int, C.POINTER(Options), C.POINTER(C.c_void_p)]
    lib.lqt_create_error.restype = C.c_char_p
    lib.lqt_destroy.argtypes = [C.c_void_p]
    lib.lqt_destroy.restype = None
    lib.lqt_last_error.argtypes = [C.c_void_p]
    lib.lqt_last_error.restype = C.c_char_p
    lib.lqt_get_info.argtypes = [C.c_void_p, C.POINTER(Info)]
    lib.lqt_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    lib.lqt_reset_stats.argtypes = [C.c_void_p]
    P, I32, I64 = C.c_void_p, C.c_int32, C.c_int64
    lib.lqt_text_project.argtypes = [P, P, I32, P]
    lib.lqt_codec_embed.argtypes = [P, P, I32, P]
    lib.lqt_code_predictor_embed.argtypes = [P, I64, I64, P]
    lib.lqt_talker_prefill.argtypes = [P, I32, P, I32, P, P]
    lib.lqt_talker_decode.argtypes = [P, I32, P, P, P]
    lib.lqt_kv_reset.argtypes = [P, I32]
    lib.lqt_kv_len.argtypes = [P, I32]
    lib.lqt_code_predictor.argtypes = [P, P, I32, I64, P]
    lib.lqt_vocoder_decode.argtypes = [P, P, I32, P, C.POINTER(I64)]
    lib.lqt_speaker_encoder.argtypes = [P, P, I32, P]
    lib.lqt_sample.argtypes = [P, P, I32, C.POINTER(Sampling), C.c_uint32, C.c_uint32, I32, C.POINTER(I64)]
    lib.lqt_generate.argtypes = [P, I32, P, I32, P, I32, P, C.POINTER(Sampling), P, I32, P,
                                 C.POINTER(I32), P, I32]
    lib.lqt_synthesize_tokens.argtypes = [P, P, I32, I32, P, C.POINTER(Sampling), P, I64,
                                          C.POINTER(I64), P, C.POINTER(I32)]
    lib.lqt_build_prompt.argtypes = [P, P, I32, I32, P, P, C.POINTER(I32), P, C.POINTER(I32), P]
    lib.lqt_debug_timeline.argtypes = [P, I32, I32, P, I32]
    lib.lqt_debug_exchange.argtypes = [P, I32, P, I32]
    lib.lqt_check_model_file.argtypes = [C.c_char_p, C.c_char_p, I32]
    lib.lqt_synthesize_batch.argtypes = [P, C.POINTER(BatchRequest), I32, C.POINTER(Sampling), C.POINTER(BatchOptions)]
    lib.lqt_debug_tc_gemm.argtypes = [P, P, P, I32, I32, I32, I32, I32, P]
    lib.lqt_synthesize_stream.argtypes = [P, P, I32, I32, P, C.POINTER(Sampling), P, I64, C.POINTER(I64), P, C.POINTER(I32), AUDIO_CB, P]
    lib.lqt_vocoder_stream_reset.argtypes = [P]
    lib.lqt_vocoder_stream_chunk.argtypes = [P, P, I32, P, C.POINTER(I64)]
    lib.lqt_log_mel.argtypes = [P, P, I64, P, C.POINTER(I32)]
    lib.lqt_speaker_embed_audio.argtypes = [P, P, I64, P]
    _lib = lib
    return lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# src/tts_onnx.h:39-47 special text-token ids
IM_START, IM_END, ASSISTANT, TTS_BOS, TTS_EOS, TTS_PAD = 151644, 151645, 77091, 151672, 151673, 151671


def wrap_text_ids(text_ids):
    """TTSEngine::synthesize id wrapping, src/tts_onnx.cpp:243-259."""
    return [IM_START, ASSISTANT, TTS_BOS] + [int(t) for t in text_ids] + [TTS_EOS, IM_END]


class EngineError(RuntimeError):
    pass


def check_model_file(path: str) -> str:
    """Header-only validation of one .lqw file (no GPU): '' = well formed, else the reason."""
    buf = C.create_string_buffer(512)
    load_library().lqt_check_model_file(path.encode(), buf, 512)
    return buf.value.decode()


class Engine:
    """One engine = one GPU = one host thread at a time (same contract as the reference)."""

    def __init__(self, model_dir: str, device: int = 0, kv_dtype: str = "bf16", n_slots: int = 0,
                 frame_impl: str = "persistent"):
        self.lib = load_library()
        h = C.c_void_p()
        opt = Options({"bf16": 0, "f32": 1}[kv_dtype], n_slots, FRAME_IMPL[frame_impl])
        rc = self.lib.lqt_create_ex(model_dir.encode(), device, C.byref(opt), C.byref(h))
        if rc != 0 or not h:
            raise EngineError(self.lib.lqt_create_error().decode())
        self.h = h
        self.info = Info()
        self.lib.lqt_get_info(self.h, C.byref(self.info))
        self.H = self.info.hidden

    def close(self):
        if getattr(self, "h", None):
            self.lib.lqt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise EngineError(self.lib.lqt_last_error(self.h).decode())

    def stats(self) -> Stats:
        s = Stats()
        self.lib.lqt_get_stats(self.h, C.byref(s))
        return s

    def reset_stats(self):
        self.lib.lqt_reset_stats(self.h)

    @staticmethod
    def sampling(temperature=0.8, top_k=50, top_p=0.95, max_new_tokens=2048, seed=0,
                 utterance_id=0, greedy=False) -> Sampling:
        return Sampling(temperature, top_p, top_k, max_new_tokens, seed, utterance_id, 1 if greedy else 0)

    # ---- per-graph runners (tts_onnx.h:196-212) ------------------------------------------------
    def text_project(self, ids):
        ids = _i64(ids)
        out = np.empty((len(ids), self.H), np.float32)
        self._ck(self.lib.lqt_text_project(self.h, _ptr(ids), len(ids), _ptr(out)))
        return out

    def codec_embed(self, ids):
        ids = _i64(ids)
        out = np.empty((len(ids), self.H), np.float32)
        self._ck(self.lib.lqt_codec_embed(self.h, _ptr(ids), len(ids), _ptr(out)))
        return out

    def code_predictor_embed(self, token: int, step: int):
        out = np.empty(self.H, np.float32)
        self._ck(self.lib.lqt_code_predictor_embed(self.h, int(token), int(step), _ptr(out)))
        return out

    def talker_prefill(self, embeds, slot: int = 0):
        e = _f32(embeds).reshape(-1, self.H)
        logits = np.empty(self.info.vocab, np.float32)
        hid = np.empty(self.H, np.float32)
        self._ck(self.lib.lqt_talker_prefill(self.h, slot, _ptr(e), e.shape[0], _ptr(logits), _ptr(hid)))
        return logits, hid

    def talker_decode(self, embed, slot: int = 0):
        e = _f32(embed).reshape(self.H)
        logits = np.empty(self.info.vocab, np.float32)
        hid = np.empty(self.H, np.float32)
        self._ck(self.lib.lqt_talker_decode(self.h, slot, _ptr(e), _ptr(logits), _ptr(hid)))
        return logits, hid

    def kv_reset(self, slot: int = 0):
        self._ck(self.lib.lqt_kv_reset(self.h, slot))

    def kv_len(self, slot: int = 0) -> int:
        return self.lib.lqt_kv_len(self.h, slot)

    def code_predictor(self, embeds, step: int):
        e = _f32(embeds).reshape(-1, self.H)
        logits = np.empty(self.info.cp_vocab, np.float32)
        self._ck(self.lib.lqt_code_predictor(self.h, _ptr(e), e.shape[0], int(step), _ptr(logits)))
        return logits

    def vocoder_decode(self, codes):
        c = _i64(codes).reshape(-1, 16)
        audio = np.empty(c.shape[0] * self.info.samples_per_frame, np.float32)
        n = C.c_int64(0)
        self._ck(self.lib.lqt_vocoder_decode(self.h, _ptr(c), c.shape[0], _ptr(audio), C.byref(n)))
        return audio[: n.value]

    def vocoder_stream(self, codes, chunks):
        """decode `codes` [T,16] in consecutive chunks of the given sizes (streaming vocoder); -> concatenated audio"""
        c = _i64(codes).reshape(-1, 16)
        self._ck(self.lib.lqt_vocoder_stream_reset(self.h))
        out, t = [], 0
        for n in chunks:
            n = min(int(n), c.shape[0] - t)
            if n <= 0:
                break
            part = np.ascontiguousarray(c[t:t + n])
            audio = np.empty(n * self.info.samples_per_frame, np.float32)
            ln = C.c_int64(0)
            self._ck(self.lib.lqt_vocoder_stream_chunk(self.h, _ptr(part), n, _ptr(audio), C.byref(ln)))
            out.append(audio[: ln.value])
            t += n
        assert t == c.shape[0], "chunk sizes do not cover the codes"
        return np.concatenate(out)

    def speaker_encoder(self, mel_t):
        m = _f32(mel_t).reshape(-1, 128)
        out = np.empty(self.H, np.float32)
        self._ck(self.lib.lqt_speaker_encoder(self.h, _ptr(m), m.shape[0], _ptr(out)))
        return out

    def log_mel(self, audio):
        """24 kHz mono f32 -> log-mel [128, frames] (reference layout) computed on the device"""
        a = _f32(audio).reshape(-1)
        nf = max(1, (a.shape[0] - 1024) // 256 + 1) if a.shape[0] >= 1024 else 1
        out = np.empty((128, nf), np.float32)
        n = C.c_int32(0)
        self._ck(self.lib.lqt_log_mel(self.h, _ptr(a), a.shape[0], _ptr(out), C.byref(n)))
        assert n.value == nf
        return out

    def speaker_embed_audio(self, audio):
        a = _f32(audio).reshape(-1)
        out = np.empty(self.H, np.float32)
        self._ck(self.lib.lqt_speaker_embed_audio(self.h, _ptr(a), a.shape[0], _ptr(out)))
        return out

    def sample(self, logits, sp: Sampling, frame: int, codebook: int, mask_codec_specials=False) -> int:
        lg = _f32(logits)
        out = C.c_int64(0)
        self._ck(self.lib.lqt_sample(self.h, _ptr(lg), lg.shape[0], C.byref(sp), frame, codebook,
                                     1 if mask_codec_specials else 0, C.byref(out)))
        return out.value

    # ---- fast path ---------------------------------------------------------------------------------
    def generate(self, prompt, trailing, tts_pad, sp: Sampling, slot: int = 0, forced_codes=None,
                 trace: bool = False):
        pr = _f32(prompt).reshape(-1, self.H)
        tr = _f32(trailing).reshape(-1, self.H)
        pad = _f32(tts_pad).reshape(self.H)
        codes = np.zeros((max(sp.max_new_tokens, 1), 16), np.int64)
        n = C.c_int32(0)
        fc = _i64(forced_codes).reshape(-1, 16) if forced_codes is not None else None
        stride = max(self.info.vocab, self.info.cp_vocab)
        tb = np.zeros((max(sp.max_new_tokens, 1), 16, stride), np.float32) if trace else None
        self._ck(self.lib.lqt_generate(self.h, slot, _ptr(pr), pr.shape[0], _ptr(tr), tr.shape[0], _ptr(pad),
                                       C.byref(sp), _ptr(fc), 0 if fc is None else fc.shape[0],
                                       _ptr(codes), C.byref(n), _ptr(tb), stride))
        return (codes[: n.value].copy(), tb) if trace else codes[: n.value].copy()

    def debug_exchange(self, which: int, n: int):
        """values of one exchange buffer of the frame kernel after the last launch (0 x, 1 qkv, 2 x1, 3 act; +4 = code predictor)"""
        out = np.empty(n, np.float32)
        self._ck(self.lib.lqt_debug_exchange(self.h, which, _ptr(out), n))
        return out

    def timeline_arm(self, entries: int = 200000, cta: int = 0):
        if self.lib.lqt_debug_timeline(self.h, entries, cta, None, 0) != 0:
            raise EngineError(self.lib.lqt_last_error(self.h).decode())
        self._tl_cap = entries

    def timeline_read(self, half: int = 0):
        """-> (clock uint64 [n], tag int [n]) of the last frame-kernel launch; half 1 = the second recorder of a profiling build
        (another consumer warp, or the producer in FK_FINE_MARKS builds)"""
        out = np.zeros(self._tl_cap, np.uint64)
        n = self.lib.lqt_debug_timeline(self.h, 0, 1 if half else 0, _ptr(out), self._tl_cap)
        out = out[: max(n, 0)]
        return (out >> np.uint64(16)), (out & np.uint64(0xFFFF)).astype(np.int64)

    def build_prompt(self, token_ids, lang: str = "auto", speaker_embed=None):
        ids = _i64(token_ids)
        prompt = np.empty((16, self.H), np.float32)
        trailing = np.empty((max(len(ids), 1), self.H), np.float32)
        pad = np.empty(self.H, np.float32)
        P, TL = C.c_int32(0), C.c_int32(0)
        spk = _f32(speaker_embed) if speaker_embed is not None else None
        self._ck(self.lib.lqt_build_prompt(self.h, _ptr(ids), len(ids), LANG_CODEC_ID[lang], _ptr(spk),
                                           _ptr(prompt), C.byref(P), _ptr(trailing), C.byref(TL), _ptr(pad)))
        return prompt[: P.value].copy(), trailing[: TL.value].copy(), pad

    def synthesize_tokens(self, token_ids, lang: str = "auto", temperature=0.8, top_k=50, top_p=0.95,
                          max_new_tokens=2048, seed=0, utterance_id=0, greedy=False, speaker_embed=None,
                          audio_out=None, codes_out=None):
        """src/tts_onnx.cpp:405-436 -> (audio f32 [n], codes i64 [T,16]).
        audio_out / codes_out: optional caller-owned (e.g. pinned) host buffers; views are returned."""
        ids = _i64(token_ids)
        sp = self.sampling(temperature, top_k, top_p, max_new_tokens, seed, utterance_id, greedy)
        cap = max(max_new_tokens, 1) * self.info.samples_per_frame
        if audio_out is not None:
            assert audio_out.dtype == np.float32 and audio_out.size >= cap and audio_out.flags.c_contiguous
            audio = audio_out.reshape(-1)
        else:
            audio = np.empty(cap, np.float32)
        if codes_out is not None:
            assert codes_out.dtype == np.int64 and codes_out.size >= max(max_new_tokens, 1) * 16
            codes = codes_out.reshape(-1)[: max(max_new_tokens, 1) * 16].reshape(-1, 16)
        else:
            codes = np.zeros((max(max_new_tokens, 1), 16), np.int64)
        ns, nf = C.c_int64(0), C.c_int32(0)
        spk = _f32(speaker_embed) if speaker_embed is not None else None
        self._ck(self.lib.lqt_synthesize_tokens(self.h, _ptr(ids), len(ids), LANG_CODEC_ID[lang], _ptr(spk),
                                                C.byref(sp), _ptr(audio), cap, C.byref(ns), _ptr(codes),
                                                C.byref(nf)))
        if audio_out is not None or codes_out is not None:
            return audio[: ns.value], codes[: nf.value]
        return audio[: ns.value].copy(), codes[: nf.value].copy()

    # ---- batched path (BASELINE configs[3], [4]) -----------------------------------------------------
    def synthesize_batch(self, requests, temperature=0.8, top_k=50, top_p=0.95, seed=0, greedy=False,
                         max_concurrent=0, planes=3, poll_frames=0, vocode=True, trace=False, pinned=None):
        """requests: list of dicts {token_ids, lang='auto', speaker_embed=None, utterance_id=i, max_new_tokens,
        forced_codes=None}. -> list of (audio f32 [n] or None, codes i64 [T,16][, trace]) in request order.
        pinned: optional list of (audio_buffer, codes_buffer) caller-owned host arrays per request."""
        n = len(requests)
        arr = (BatchRequest * n)()
        keep, outs = [], []
        stride = max(self.info.vocab, self.info.cp_vocab)
        for i, r in enumerate(requests):
            ids = _i64(r["token_ids"])
            mx = int(r["max_new_tokens"])
            spk = _f32(r["speaker_embed"]) if r.get("speaker_embed") is not None else None
            fc = _i64(r["forced_codes"]).reshape(-1, 16) if r.get("forced_codes") is not None else None
            if pinned is not None:
                audio, codes = pinned[i]
            else:
                audio = np.empty(max(mx, 1) * self.info.samples_per_frame, np.float32) if vocode else None
                codes = np.zeros((max(mx, 1), 16), np.int64)
            tb = np.zeros((max(mx, 1), 16, stride), np.float32) if trace else None
            ns, nf = C.c_int64(0), C.c_int32(0)
            keep.append((ids, spk, fc, audio, codes, tb, ns, nf))
            a = arr[i]
            a.token_ids = _ptr(ids); a.n_ids = len(ids); a.lang_codec_id = LANG_CODEC_ID[r.get("lang", "auto")]
            a.speaker_embed = _ptr(spk); a.utterance_id = int(r.get("utterance_id", i)); a.max_new_tokens = mx
            a.forced_codes = _ptr(fc); a.n_forced = 0 if fc is None else fc.shape[0]
            a.audio_out = _ptr(audio) if vocode else None
            a.audio_capacity = 0 if audio is None else audio.size
            a.n_samples = C.pointer(ns); a.codes_out = _ptr(codes); a.n_frames = C.pointer(nf)
            a.logits_trace = _ptr(tb)
        sp = self.sampling(temperature, top_k, top_p, 0, seed, 0, greedy)
        opt = BatchOptions(max_concurrent, planes, poll_frames)
        self._ck(self.lib.lqt_synthesize_batch(self.h, arr, n, C.byref(sp), C.byref(opt)))
        for ids, spk, fc, audio, codes, tb, ns, nf in keep:
            a = audio.reshape(-1)[: ns.value] if (audio is not None and vocode) else None
            c = codes.reshape(-1, 16)[: nf.value]
            outs.append((a, c, tb) if trace else (a, c))
        return outs

    def debug_tc_gemm(self, W, x, planes=3, splits=0):
        """out[b][n] = sum_k bf16(W[n][k]) * x[b][k] through the tcgen05 GEMM of the batched path (parity surface)"""
        W, x = _f32(W), _f32(x)
        out = np.empty((x.shape[0], W.shape[0]), np.float32)
        self._ck(self.lib.lqt_debug_tc_gemm(self.h, _ptr(W), _ptr(x), W.shape[0], W.shape[1], x.shape[0], planes, splits, _ptr(out)))
        return out

    def synthesize_stream(self, token_ids, lang="auto", temperature=0.8, top_k=50, top_p=0.95, max_new_tokens=2048, seed=0,
                          utterance_id=0, on_chunk=None):
        """lqt_synthesize_stream: on_chunk(first_sample, pcm ndarray copy, seconds since the call started) per PCM chunk, in order,
        while generation continues. -> (audio, codes)"""
        import time
        ids = _i64(token_ids)
        sp = self.sampling(temperature, top_k, top_p, max_new_tokens, seed, utterance_id, False)
        cap = max(max_new_tokens, 1) * self.info.samples_per_frame
        audio = np.empty(cap, np.float32)
        codes = np.zeros((max(max_new_tokens, 1), 16), np.int64)
        ns, nf = C.c_int64(0), C.c_int32(0)
        t0 = time.perf_counter()

        def _cb(user, pcm, first, n):
            if on_chunk is not None:
                on_chunk(int(first), np.ctypeslib.as_array(pcm, shape=(int(n),)).copy(), time.perf_counter() - t0)
        cb = AUDIO_CB(_cb)
        self._ck(self.lib.lqt_synthesize_stream(self.h, _ptr(ids), len(ids), LANG_CODEC_ID[lang], None, C.byref(sp), _ptr(audio), cap,
                                                C.byref(ns), _ptr(codes), C.byref(nf), cb, None))
        return audio[: ns.value].copy(), codes[: nf.value].copy()
