"""Frozen model spec, weight-file format (.lqw) and seeded random-init generator.

The reference (leaxer-ai/leaxer-qwen3-tts) pins only the *I/O contract* of its seven ONNX graphs
(/root/reference/src/tts_onnx.cpp:545-776) and a handful of dimensions (src/tts_onnx.h:29-70).
The graphs themselves are not vendored, so the internal architecture is frozen HERE (SURVEY.md §8
header) and shared bit-for-bit by the oracle (oracle/) and the CUDA engine (csrc/).

Model directory layout mirrors the reference's seven file names (src/tts_onnx.cpp:91-107), with the
extension `.lqw` instead of `.onnx` (ONNX cannot be produced or parsed in this environment):

    text_project.lqw  codec_embed.lqw  code_predictor_embed.lqw  talker_prefill.lqw
    talker_decode.lqw (shares talker_prefill's tensors)  code_predictor.lqw
    tokenizer12hz_decode.lqw  [speaker_encoder.lqw]

.lqw file format (little endian):
    8 bytes   magic "LQTW0001"
    u32       n_tensors
    u32       n_meta
    u64       data_start (file offset of the data section, 256-byte aligned)
    n_meta  x { u16 klen, key bytes, u16 vlen, value bytes }          (model-spec key/values, text)
    n_tensors x { u16 nlen, name bytes, u8 dtype (0=bf16, 1=f32), u8 ndim, u32 dims[ndim],
                  u64 offset (relative to data_start, 256-byte aligned), u64 nbytes }
    data section

All matrices are stored "GEMM ready": Linear weights [N_out, K_in] row-major bf16; conv weights
[C_out, taps, C_in]; transposed-conv weights [stride, C_out, 2, C_in] (phase-major, see vocoder).
Vectors (norm weights, biases, snake/scale parameters) are f32.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass, field, asdict

import numpy as np

MAGIC = b"LQTW0001"
DT_BF16, DT_F32 = 0, 1

GRAPH_FILES = [
    "text_project", "codec_embed", "code_predictor_embed", "talker_prefill",
    "talker_decode", "code_predictor", "tokenizer12hz_decode",
]
OPTIONAL_GRAPH_FILES = ["speaker_encoder"]


# ----------------------------------------------------------------------------------------------
# Spec
# ----------------------------------------------------------------------------------------------
@dataclass
class ModelSpec:
    name: str = "qwen3-tts-0.6b"
    # talker (src/tts_onnx.h:31-35 pins hidden/layers/kv_heads/head_dim/vocab for 0.6B)
    hidden: int = 1024
    layers: int = 28
    heads: int = 16
    kv_heads: int = 8
    head_dim: int = 128
    inter: int = 3072
    vocab: int = 3072
    rope_theta: float = 1e6
    rms_eps: float = 1e-6
    max_pos: int = 2304            # >= P(10) + MAX_NEW_TOKENS(2048)
    # code predictor (always `cp_hidden` wide; 1.7B talker feeds it through in_proj)
    cp_hidden: int = 1024
    cp_layers: int = 5
    cp_heads: int = 16
    cp_kv_heads: int = 8
    cp_inter: int = 3072
    cp_vocab: int = 2048           # src/tts_onnx.h:37
    cp_steps: int = 15             # NUM_CODE_GROUPS - 1
    cp_max_pos: int = 32
    # text_project
    text_vocab: int = 151936
    text_dim: int = 2048
    # vocoder (tokenizer12hz_decode)
    voc_codebook_size: int = 2048
    voc_codebook_dim: int = 256
    voc_rvq_out: int = 512
    voc_hidden: int = 1024
    voc_layers: int = 8
    voc_heads: int = 16
    voc_head_dim: int = 64
    voc_inter: int = 3072
    voc_window: int = 72
    voc_rope_theta: float = 1e4
    voc_rms_eps: float = 1e-5
    voc_max_pos: int = 2304
    voc_upsampling_ratios: tuple = (2, 2)
    voc_decoder_dim: int = 1536
    voc_upsample_rates: tuple = (8, 5, 4, 3)
    # speaker encoder (optional 8th graph, src/tts_onnx.cpp:367-403)
    spk_mels: int = 128
    spk_channels: int = 512
    spk_layers: int = 3
    seed: int = 0

    @property
    def q_dim(self): return self.heads * self.head_dim
    @property
    def kv_dim(self): return self.kv_heads * self.head_dim
    @property
    def qkv_dim(self): return self.q_dim + 2 * self.kv_dim
    @property
    def cp_q_dim(self): return self.cp_heads * self.head_dim
    @property
    def cp_kv_dim(self): return self.cp_kv_heads * self.head_dim
    @property
    def samples_per_frame(self):
        n = 1
        for r in tuple(self.voc_upsampling_ratios) + tuple(self.voc_upsample_rates):
            n *= r
        return n

    def to_meta(self) -> dict:
        d = asdict(self)
        out = {}
        for k, v in d.items():
            if isinstance(v, (tuple, list)):
                out[k] = ",".join(str(x) for x in v)
            else:
                out[k] = repr(v) if isinstance(v, float) else str(v)
        return out

    @staticmethod
    def from_meta(meta: dict) -> "ModelSpec":
        s = ModelSpec()
        for k, v in meta.items():
            if not hasattr(s, k):
                continue
            cur = getattr(s, k)
            if isinstance(cur, tuple):
                setattr(s, k, tuple(int(x) for x in v.split(",") if x))
            elif isinstance(cur, bool):
                setattr(s, k, v == "True")
            elif isinstance(cur, int):
                setattr(s, k, int(v))
            elif isinstance(cur, float):
                setattr(s, k, float(v))
            else:
                setattr(s, k, v)
        return s


def spec_0p6b(seed: int = 0) -> ModelSpec:
    return ModelSpec(seed=seed)


def spec_1p7b(seed: int = 0) -> ModelSpec:
    """BASELINE.json config 5: 1.7B talker (hidden 2048, inter 6144); predictor stays 1024 wide."""
    return ModelSpec(name="qwen3-tts-1.7b", hidden=2048, inter=6144, seed=seed)


def spec_tiny(seed: int = 0) -> ModelSpec:
    """Small spec with the same structure, for fast CPU tests of oracle/host logic and for
    GPU kernel tests at non-default dims. Token ids keep their real values."""
    return ModelSpec(
        name="qwen3-tts-tiny", hidden=256, layers=2, heads=4, kv_heads=2, inter=512,
        cp_hidden=256, cp_layers=2, cp_heads=4, cp_kv_heads=2, cp_inter=512,
        text_dim=64, max_pos=256, voc_max_pos=256,
        voc_codebook_dim=32, voc_rvq_out=64, voc_hidden=128, voc_layers=2, voc_heads=2,
        voc_inter=256, voc_window=8, voc_decoder_dim=192, spk_channels=64, seed=seed)


# ----------------------------------------------------------------------------------------------
# Deterministic generator (pure uint32 / float32 arithmetic: identical on every machine)
# ----------------------------------------------------------------------------------------------
def _mix32(x: np.ndarray) -> np.ndarray:
    """lowbias32 integer hash, vectorised, in place on a uint32 array."""
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def _mix32_scalar(x: int) -> int:
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def _name_seed(seed: int, name: str) -> int:
    h = 0x811C9DC5
    for b in name.encode():
        h = ((h ^ b) * 0x01000193) & 0xFFFFFFFF
    return _mix32_scalar(h ^ _mix32_scalar(seed + 0x9E3779B9))


def uniform_pm1(seed: int, name: str, n: int, start: int = 0) -> np.ndarray:
    """n float32 values in (-1, 1): u23 = hash(idx) >> 9 ; v = (u23 + 0.5) * 2^-22 - 1 (all exact)."""
    ts = np.uint32(_name_seed(seed, name))
    idx = np.arange(start, start + n, dtype=np.uint32)
    idx ^= ts
    _mix32(idx)
    idx >>= np.uint32(9)
    v = idx.astype(np.float32)
    v += np.float32(0.5)
    v *= np.float32(2.0 ** -22)
    v -= np.float32(1.0)
    return v


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even float32 -> bf16 bit pattern (uint16). Finite inputs only."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = b + np.uint32(0x7FFF) + ((b >> np.uint32(16)) & np.uint32(1))
    return (r >> np.uint32(16)).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << np.uint32(16)).view(np.float32)


@dataclass
class TensorDef:
    name: str
    shape: tuple
    dtype: int            # DT_BF16 / DT_F32
    scale: float = 0.02   # value = offset + scale * uniform(-1, 1)
    offset: float = 0.0
    kind: str = "rand"    # "rand" | "rope_cos" | "rope_sin" | "zero_row"
    extra: dict = field(default_factory=dict)


_S3 = 3.0 ** 0.5  # uniform(-a, a) has std a/sqrt(3)


def _layer_defs(prefix: str, H: int, qd: int, kvd: int, D: int, I: int, qk_norm: bool,
                layer_scale: bool) -> list:
    w = 0.02 * _S3
    d = [
        TensorDef(f"{prefix}.ln1", (H,), DT_F32, 0.1, 1.0),
        TensorDef(f"{prefix}.wqkv", (qd + 2 * kvd, H), DT_BF16, w),
        TensorDef(f"{prefix}.wo", (H, qd), DT_BF16, w),
        TensorDef(f"{prefix}.ln2", (H,), DT_F32, 0.1, 1.0),
        TensorDef(f"{prefix}.wgate", (I, H), DT_BF16, w),
        TensorDef(f"{prefix}.wup", (I, H), DT_BF16, w),
        TensorDef(f"{prefix}.wdown", (H, I), DT_BF16, w),
    ]
    if qk_norm:
        d += [TensorDef(f"{prefix}.qnorm", (D,), DT_F32, 0.1, 1.0),
              TensorDef(f"{prefix}.knorm", (D,), DT_F32, 0.1, 1.0)]
    if layer_scale:
        d += [TensorDef(f"{prefix}.ls1", (H,), DT_F32, 0.05, 0.10),
              TensorDef(f"{prefix}.ls2", (H,), DT_F32, 0.05, 0.10)]
    return d


def graph_tensor_defs(spec: ModelSpec) -> dict:
    """name of graph file -> list[TensorDef]. This IS the architecture spec."""
    H, D = spec.hidden, spec.head_dim
    g = {}
    # text_project: embedding -> Linear+b -> SiLU -> Linear+b  (SURVEY §8 header)
    g["text_project"] = [
        TensorDef("embed", (spec.text_vocab, spec.text_dim), DT_BF16, 1.0),
        TensorDef("fc1.weight", (spec.text_dim, spec.text_dim), DT_BF16, 0.03 * _S3),
        TensorDef("fc1.bias", (spec.text_dim,), DT_F32, 0.1),
        TensorDef("fc2.weight", (H, spec.text_dim), DT_BF16, 0.05 * _S3),
        TensorDef("fc2.bias", (H,), DT_F32, 0.1),
    ]
    g["codec_embed"] = [TensorDef("embed", (spec.vocab, H), DT_BF16, 1.0)]
    g["code_predictor_embed"] = [TensorDef("embed", (spec.cp_steps, spec.cp_vocab, H), DT_BF16, 1.0)]
    # talker (prefill and decode share one weight set, like the two ONNX exports of one model)
    t = []
    for i in range(spec.layers):
        t += _layer_defs(f"l{i}", H, spec.q_dim, spec.kv_dim, D, spec.inter, True, False)
    t += [TensorDef("norm", (H,), DT_F32, 0.1, 1.0),
          # EOS row (2150) is zero so random-init runs never stop early (SURVEY §7 hard parts)
          TensorDef("head", (spec.vocab, H), DT_BF16, 0.1 * _S3, kind="zero_row",
                    extra={"rows": [2150]}),
          TensorDef("rope_cos", (spec.max_pos, D // 2), DT_F32, kind="rope_cos",
                    extra={"theta": spec.rope_theta}),
          TensorDef("rope_sin", (spec.max_pos, D // 2), DT_F32, kind="rope_sin",
                    extra={"theta": spec.rope_theta})]
    g["talker_prefill"] = t
    g["talker_decode"] = []        # shares talker_prefill's tensors (meta: shares=talker_prefill)
    # code predictor
    Hc = spec.cp_hidden
    c = []
    if spec.hidden != Hc:
        c += [TensorDef("in_proj.weight", (Hc, H), DT_BF16, 0.03 * _S3),
              TensorDef("in_proj.bias", (Hc,), DT_F32, 0.1)]
    for i in range(spec.cp_layers):
        c += _layer_defs(f"l{i}", Hc, spec.cp_q_dim, spec.cp_kv_dim, D, spec.cp_inter, True, False)
    c += [TensorDef("norm", (Hc,), DT_F32, 0.1, 1.0),
          TensorDef("heads", (spec.cp_steps, spec.cp_vocab, Hc), DT_BF16, 0.1 * _S3),
          TensorDef("rope_cos", (spec.cp_max_pos, D // 2), DT_F32, kind="rope_cos",
                    extra={"theta": spec.rope_theta}),
          TensorDef("rope_sin", (spec.cp_max_pos, D // 2), DT_F32, kind="rope_sin",
                    extra={"theta": spec.rope_theta})]
    g["code_predictor"] = c
    # vocoder
    Cv, Dc, R = spec.voc_hidden, spec.voc_codebook_dim, spec.voc_rvq_out
    v = [
        TensorDef("rvq.sem.codebook", (1, spec.voc_codebook_size, Dc), DT_BF16, 1.0),
        TensorDef("rvq.sem.out_proj", (R, Dc), DT_BF16, (1.0 / Dc ** 0.5) * _S3),
        TensorDef("rvq.aco.codebook", (spec.cp_steps, spec.voc_codebook_size, Dc), DT_BF16, 1.0),
        TensorDef("rvq.aco.out_proj", (R, Dc), DT_BF16, (0.25 / Dc ** 0.5) * _S3),
        TensorDef("pre_conv.weight", (Cv, 3, R), DT_BF16, (1.0 / (3 * R) ** 0.5) * _S3),
        TensorDef("pre_conv.bias", (Cv,), DT_F32, 0.1),
    ]
    vqd = spec.voc_heads * spec.voc_head_dim
    for i in range(spec.voc_layers):
        v += _layer_defs(f"pt.l{i}", Cv, vqd, vqd, spec.voc_head_dim, spec.voc_inter, False, True)
    v += [TensorDef("pt.norm", (Cv,), DT_F32, 0.1, 1.0),
          TensorDef("pt.rope_cos", (spec.voc_max_pos, spec.voc_head_dim // 2), DT_F32,
                    kind="rope_cos", extra={"theta": spec.voc_rope_theta}),
          TensorDef("pt.rope_sin", (spec.voc_max_pos, spec.voc_head_dim // 2), DT_F32,
                    kind="rope_sin", extra={"theta": spec.voc_rope_theta})]
    for u, f in enumerate(spec.voc_upsampling_ratios):
        # transposed conv kernel = stride = f: [phase f][C_out][1][C_in]
        v += [TensorDef(f"up{u}.tconv.weight", (f, Cv, 1, Cv), DT_BF16, (1.0 / Cv ** 0.5) * _S3),
              TensorDef(f"up{u}.tconv.bias", (Cv,), DT_F32, 0.1),
              TensorDef(f"up{u}.dw.weight", (7, Cv), DT_F32, (1.0 / 7 ** 0.5) * _S3),
              TensorDef(f"up{u}.dw.bias", (Cv,), DT_F32, 0.1),
              TensorDef(f"up{u}.ln.weight", (Cv,), DT_F32, 0.1, 1.0),
              TensorDef(f"up{u}.ln.bias", (Cv,), DT_F32, 0.1),
              TensorDef(f"up{u}.pw1.weight", (4 * Cv, Cv), DT_BF16, (1.0 / Cv ** 0.5) * _S3),
              TensorDef(f"up{u}.pw1.bias", (4 * Cv,), DT_F32, 0.1),
              TensorDef(f"up{u}.pw2.weight", (Cv, 4 * Cv), DT_BF16, (1.0 / (4 * Cv) ** 0.5) * _S3),
              TensorDef(f"up{u}.pw2.bias", (Cv,), DT_F32, 0.1),
              TensorDef(f"up{u}.gamma", (Cv,), DT_F32, 0.05, 0.10)]
    Cd = spec.voc_decoder_dim
    v += [TensorDef("dec.conv_in.weight", (Cd, 7, Cv), DT_BF16, (1.0 / (7 * Cv) ** 0.5) * _S3),
          TensorDef("dec.conv_in.bias", (Cd,), DT_F32, 0.1)]
    cin = Cd
    for b, s in enumerate(spec.voc_upsample_rates):
        cout = cin // 2
        v += [TensorDef(f"dec.b{b}.snake.alpha", (cin,), DT_F32, 0.3),
              TensorDef(f"dec.b{b}.snake.beta", (cin,), DT_F32, 0.3),
              # transposed conv kernel 2s stride s: [phase s][C_out][2 (x[p], x[p-1])][C_in]
              TensorDef(f"dec.b{b}.tconv.weight", (s, cout, 2, cin), DT_BF16,
                        (1.0 / (2 * cin) ** 0.5) * _S3),
              TensorDef(f"dec.b{b}.tconv.bias", (cout,), DT_F32, 0.1)]
        for r in range(3):
            p = f"dec.b{b}.r{r}"
            v += [TensorDef(f"{p}.snake1.alpha", (cout,), DT_F32, 0.3),
                  TensorDef(f"{p}.snake1.beta", (cout,), DT_F32, 0.3),
                  TensorDef(f"{p}.conv1.weight", (cout, 7, cout), DT_BF16,
                            (0.7 / (7 * cout) ** 0.5) * _S3),
                  TensorDef(f"{p}.conv1.bias", (cout,), DT_F32, 0.05),
                  TensorDef(f"{p}.snake2.alpha", (cout,), DT_F32, 0.3),
                  TensorDef(f"{p}.snake2.beta", (cout,), DT_F32, 0.3),
                  TensorDef(f"{p}.conv2.weight", (cout, 1, cout), DT_BF16,
                            (0.5 / cout ** 0.5) * _S3),
                  TensorDef(f"{p}.conv2.bias", (cout,), DT_F32, 0.05)]
        cin = cout
    v += [TensorDef("dec.snake_out.alpha", (cin,), DT_F32, 0.3),
          TensorDef("dec.snake_out.beta", (cin,), DT_F32, 0.3),
          TensorDef("dec.conv_out.weight", (1, 7, cin), DT_F32, (0.12 / (7 * cin) ** 0.5) * _S3),
          TensorDef("dec.conv_out.bias", (1,), DT_F32, 0.01)]
    g["tokenizer12hz_decode"] = v
    # speaker encoder (optional): log-mel [frames,128] -> conv k5 stack + ReLU -> mean/std pooling
    # -> Linear -> [hidden]   (stand-in for upstream's ECAPA-TDNN)
    Cs = spec.spk_channels
    s_ = [TensorDef("in_conv.weight", (Cs, 5, spec.spk_mels), DT_BF16,
                    (1.0 / (5 * spec.spk_mels) ** 0.5) * _S3 * 0.2),
          TensorDef("in_conv.bias", (Cs,), DT_F32, 0.1)]
    for i in range(spec.spk_layers):
        s_ += [TensorDef(f"l{i}.conv.weight", (Cs, 3, Cs), DT_BF16, (1.0 / (3 * Cs) ** 0.5) * _S3),
               TensorDef(f"l{i}.conv.bias", (Cs,), DT_F32, 0.1)]
    s_ += [TensorDef("fc.weight", (H, 2 * Cs), DT_BF16, (1.0 / (2 * Cs) ** 0.5) * _S3),
           TensorDef("fc.bias", (H,), DT_F32, 0.1)]
    g["speaker_encoder"] = s_
    return g


def _gen_tensor(spec: ModelSpec, gname: str, td: TensorDef) -> np.ndarray:
    """Returns uint16 (bf16 bits) or float32 array of td.shape."""
    n = int(np.prod(td.shape))
    if td.kind in ("rope_cos", "rope_sin"):
        half = td.shape[1]
        inv = 1.0 / (td.extra["theta"] ** (np.arange(half, dtype=np.float64) * 2.0 / (2 * half)))
        ang = np.arange(td.shape[0], dtype=np.float64)[:, None] * inv[None, :]
        v = (np.cos(ang) if td.kind == "rope_cos" else np.sin(ang)).astype(np.float32)
        return v
    out = np.empty(n, dtype=np.uint16 if td.dtype == DT_BF16 else np.float32)
    CH = 1 << 24
    full = f"{gname}/{td.name}"
    for s in range(0, n, CH):
        m = min(CH, n - s)
        v = uniform_pm1(spec.seed, full, m, s)
        v *= np.float32(td.scale)
        if td.offset != 0.0:
            v += np.float32(td.offset)
        out[s:s + m] = f32_to_bf16_bits(v) if td.dtype == DT_BF16 else v
    out = out.reshape(td.shape)
    if td.kind == "zero_row":
        for r in td.extra["rows"]:
            if r < td.shape[0]:
                out[r] = 0
    return out


# ----------------------------------------------------------------------------------------------
# File I/O
# ----------------------------------------------------------------------------------------------
def _align(x: int, a: int = 256) -> int:
    return (x + a - 1) // a * a


def write_lqw(path: str, tensors: list, meta: dict) -> None:
    """tensors: list of (name, dtype_code, np.ndarray)."""
    entries, off = [], 0
    for name, dt, arr in tensors:
        nb = arr.nbytes
        entries.append((name, dt, arr.shape, off, nb))
        off = _align(off + nb)
    hdr = bytearray()
    for k, v in meta.items():
        kb, vb = k.encode(), str(v).encode()
        hdr += struct.pack("<H", len(kb)) + kb + struct.pack("<H", len(vb)) + vb
    for name, dt, shape, o, nb in entries:
        nbs = name.encode()
        hdr += struct.pack("<H", len(nbs)) + nbs + struct.pack("<BB", dt, len(shape))
        hdr += struct.pack(f"<{len(shape)}I", *shape) + struct.pack("<QQ", o, nb)
    data_start = _align(8 + 4 + 4 + 8 + len(hdr))
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(MAGIC + struct.pack("<IIQ", len(entries), len(meta), data_start) + hdr)
        for (name, dt, arr), (_, _, _, o, nb) in zip(tensors, entries):
            f.seek(data_start + o)
            f.write(memoryview(np.ascontiguousarray(arr)).cast("B"))
        f.truncate(data_start + off)      # zero-pad the tail to the aligned end
    os.replace(tmp, path)


def read_lqw(path: str):
    """-> (meta dict, {name: np.ndarray}) ; bf16 tensors come back as uint16 bit patterns (mmap)."""
    with open(path, "rb") as f:
        head = f.read(24)
        if head[:8] != MAGIC:
            raise ValueError(f"{path}: bad magic")
        nt, nm, data_start = struct.unpack("<IIQ", head[8:24])
        hdr = f.read(data_start - 24)
    p = 0
    meta = {}
    for _ in range(nm):
        kl, = struct.unpack_from("<H", hdr, p); p += 2
        k = hdr[p:p + kl].decode(); p += kl
        vl, = struct.unpack_from("<H", hdr, p); p += 2
        meta[k] = hdr[p:p + vl].decode(); p += vl
    tensors = {}
    mm = np.memmap(path, dtype=np.uint8, mode="r") if nt else None
    for _ in range(nt):
        nl, = struct.unpack_from("<H", hdr, p); p += 2
        name = hdr[p:p + nl].decode(); p += nl
        dt, nd = struct.unpack_from("<BB", hdr, p); p += 2
        shape = struct.unpack_from(f"<{nd}I", hdr, p); p += 4 * nd
        o, nb = struct.unpack_from("<QQ", hdr, p); p += 16
        raw = mm[data_start + o: data_start + o + nb]
        tensors[name] = raw.view(np.uint16 if dt == DT_BF16 else np.float32).reshape(shape)
    return meta, tensors


def generate_model_dir(out_dir: str, spec: ModelSpec, with_speaker_encoder: bool = True,
                       verbose: bool = False) -> str:
    """Writes the 7(+1) .lqw files for `spec` into out_dir (idempotent: skips complete dirs)."""
    os.makedirs(out_dir, exist_ok=True)
    stamp = os.path.join(out_dir, ".complete")
    want = repr(sorted(spec.to_meta().items())) + str(with_speaker_encoder)
    if os.path.exists(stamp) and open(stamp).read() == want:
        return out_dir
    defs = graph_tensor_defs(spec)
    names = list(GRAPH_FILES) + (OPTIONAL_GRAPH_FILES if with_speaker_encoder else [])
    for gname in names:
        meta = dict(spec.to_meta())
        meta["graph"] = gname
        if gname == "talker_decode":
            meta["shares"] = "talker_prefill"
        tensors = []
        for td in defs[gname]:
            arr = _gen_tensor(spec, gname if gname != "talker_decode" else "talker_prefill", td)
            tensors.append((td.name, td.dtype, arr))
        write_lqw(os.path.join(out_dir, gname + ".lqw"), tensors, meta)
        if verbose:
            print(f"[modelspec] wrote {gname}.lqw ({sum(a.nbytes for _, _, a in tensors) / 1e6:.1f} MB)")
    with open(stamp, "w") as f:
        f.write(want)
    return out_dir


def load_model_dir(model_dir: str):
    """-> (spec, {graph: {tensor: np.ndarray}})"""
    graphs, spec = {}, None
    for gname in GRAPH_FILES + OPTIONAL_GRAPH_FILES:
        p = os.path.join(model_dir, gname + ".lqw")
        if not os.path.exists(p):
            continue
        meta, tensors = read_lqw(p)
        if spec is None:
            spec = ModelSpec.from_meta(meta)
        graphs[gname] = tensors
    if "talker_prefill" in graphs:
        graphs["talker_decode"] = graphs["talker_prefill"]
    return spec, graphs


def synthetic_text_ids(n_text: int, seed: int = 1234):
    """Synthetic prompt ids uniform in [0,151643); hash-based, machine independent.
    Same stream as oracle.synthetic_text_ids (tests assert equality)."""
    u = uniform_pm1(seed, "synthetic_text_ids", n_text)
    return [int(x) for x in ((u.astype(np.float64) + 1.0) * 0.5 * 151643).astype(np.int64)]


def default_model_dir(spec: ModelSpec) -> str:
    root = os.environ.get("LQT_MODEL_CACHE", "/tmp/lqt_models")
    return os.path.join(root, f"{spec.name}-seed{spec.seed}", "onnx_kv")


if __name__ == "__main__":
    import argparse, time
    ap = argparse.ArgumentParser(description="write a seeded random-init model directory")
    ap.add_argument("--spec", default="0.6b", choices=["0.6b", "1.7b", "tiny"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    sp = {"0.6b": spec_0p6b, "1.7b": spec_1p7b, "tiny": spec_tiny}[a.spec](a.seed)
    t0 = time.time()
    d = generate_model_dir(a.out or default_model_dir(sp), sp, verbose=True)
    print(f"model dir: {d}  ({time.time() - t0:.1f}s)")
