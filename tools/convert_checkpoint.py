#!/usr/bin/env python
"""Offline converter: a Qwen3-TTS checkpoint in HuggingFace safetensors form -> the engine's seven (+1) .lqw files.

    python tools/convert_checkpoint.py --src DIR_WITH_SAFETENSORS --out onnx_kv_06b [--spec 0.6b|1.7b] [--strict]
    python tools/convert_checkpoint.py --export-hf DIR --from-lqw MODEL_DIR        (inverse: used by the round-trip test)

Why this exists: lqt_create only reads .lqw files, and until now only modelspec.py (seeded random init) could
write them, so `TTSEngine("onnx/onnx_kv_06b")` had nothing real to load. The reference's own models are ONNX exports of the
same checkpoint (README.md:69-93, HF zukky/Qwen3-TTS-ONNX-DLL); this tool goes from the checkpoint those were exported from.

What it does: every tensor the engine needs (modelspec.graph_tensor_defs) is produced from one or more source tensors by a
declarative rule (NAME_MAP below): rename, concatenate q|k|v, re-lay convolution weights (PyTorch [Cout, Cin, k] ->
[Cout, k, Cin]; transposed conv [Cin, Cout, k] -> phase-major [stride, Cout, 2, Cin]), cast to bf16 / f32. Every result is
checked against the spec's shape and dtype before anything is written; missing or mis-shaped sources are listed and the run
fails (no partial model directory). RoPE tables are recomputed from the spec (they are not parameters).

STATUS: no checkpoint, no network and no `onnx` package exist in this environment, so the source names follow the
transformers implementation of the same architecture family (models/qwen3_omni_moe/modeling_qwen3_omni_moe.py: talker,
code predictor, Code2Wav) and are VERIFIED ONLY BY ROUND TRIP (tests/test_convert_checkpoint.py: synthetic model -> HF-named
safetensors -> .lqw, byte-identical). A real checkpoint whose names differ is reported tensor by tensor; `--map FILE.json`
overrides source names without touching the code. ONNX initialisers as a source need the `onnx` package (`--src-onnx`)."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

load_package()
from leaxer_qwen3_tts_b200 import modelspec as ms  # noqa: E402


# ----------------------------------------------------------------------------------------------
# layout transforms (forward = checkpoint -> engine; inverse = engine -> checkpoint, for the round trip)
# ----------------------------------------------------------------------------------------------
def conv_fwd(w):            # torch Conv1d [Cout, Cin, k] -> [Cout, k, Cin]
    return np.ascontiguousarray(np.transpose(w, (0, 2, 1)))


def conv_inv(w):
    return np.ascontiguousarray(np.transpose(w, (0, 2, 1)))


def tconv_fwd(w, stride):   # torch ConvTranspose1d [Cin, Cout, k = nh*stride] -> [stride, Cout, nh, Cin]; tap index k = h*stride + r
    cin, cout, k = w.shape
    nh = k // stride
    return np.ascontiguousarray(np.transpose(w.reshape(cin, cout, nh, stride), (3, 1, 2, 0)))


def tconv_inv(w):
    s, cout, nh, cin = w.shape
    return np.ascontiguousarray(np.transpose(w, (3, 1, 2, 0)).reshape(cin, cout, nh * s))


def dw_fwd(w):              # depthwise Conv1d [C, 1, 7] -> [7, C]
    return np.ascontiguousarray(w[:, 0, :].T)


def dw_inv(w):
    return np.ascontiguousarray(w.T[:, None, :])


# ----------------------------------------------------------------------------------------------
# name map: engine (graph, tensor) -> rule. A rule is (kind, source name(s)[, arg]).
# ----------------------------------------------------------------------------------------------
def _layer_rules(dst_prefix: str, src_prefix: str, qk_norm: bool, layer_scale: bool) -> dict:
    a, m = f"{src_prefix}.self_attn", f"{src_prefix}.mlp"
    r = {
        f"{dst_prefix}.ln1": ("copy", f"{src_prefix}.input_layernorm.weight"),
        f"{dst_prefix}.wqkv": ("cat0", [f"{a}.q_proj.weight", f"{a}.k_proj.weight", f"{a}.v_proj.weight"]),
        f"{dst_prefix}.wo": ("copy", f"{a}.o_proj.weight"),
        f"{dst_prefix}.ln2": ("copy", f"{src_prefix}.post_attention_layernorm.weight"),
        f"{dst_prefix}.wgate": ("copy", f"{m}.gate_proj.weight"),
        f"{dst_prefix}.wup": ("copy", f"{m}.up_proj.weight"),
        f"{dst_prefix}.wdown": ("copy", f"{m}.down_proj.weight"),
    }
    if qk_norm:
        r[f"{dst_prefix}.qnorm"] = ("copy", f"{a}.q_norm.weight")
        r[f"{dst_prefix}.knorm"] = ("copy", f"{a}.k_norm.weight")
    if layer_scale:
        r[f"{dst_prefix}.ls1"] = ("copy", f"{src_prefix}.self_attn_layer_scale.scale")
        r[f"{dst_prefix}.ls2"] = ("copy", f"{src_prefix}.mlp_layer_scale.scale")
    return r


def name_map(spec: ms.ModelSpec) -> dict:
    g: dict = {}
    g["text_project"] = {
        "embed": ("copy", "talker.model.text_embedding.weight"),
        "fc1.weight": ("copy", "talker.text_projection.linear_fc1.weight"), "fc1.bias": ("copy", "talker.text_projection.linear_fc1.bias"),
        "fc2.weight": ("copy", "talker.text_projection.linear_fc2.weight"), "fc2.bias": ("copy", "talker.text_projection.linear_fc2.bias"),
    }
    g["codec_embed"] = {"embed": ("copy", "talker.model.codec_embedding.weight")}
    g["code_predictor_embed"] = {"embed": ("stack", [f"talker.code_predictor.model.codec_embedding.{j}.weight" for j in range(spec.cp_steps)])}
    t = {}
    for i in range(spec.layers):
        t.update(_layer_rules(f"l{i}", f"talker.model.layers.{i}", True, False))
    t["norm"] = ("copy", "talker.model.norm.weight")
    t["head"] = ("copy", "talker.codec_head.weight")
    t["rope_cos"] = ("rope", None); t["rope_sin"] = ("rope", None)
    g["talker_prefill"] = t
    c = {}
    if spec.hidden != spec.cp_hidden:
        c["in_proj.weight"] = ("copy", "talker.code_predictor.small_to_mtp_projection.weight")
        c["in_proj.bias"] = ("copy", "talker.code_predictor.small_to_mtp_projection.bias")
    for i in range(spec.cp_layers):
        c.update(_layer_rules(f"l{i}", f"talker.code_predictor.model.layers.{i}", True, False))
    c["norm"] = ("copy", "talker.code_predictor.model.norm.weight")
    c["heads"] = ("stack", [f"talker.code_predictor.lm_head.{j}.weight" for j in range(spec.cp_steps)])
    c["rope_cos"] = ("rope", None); c["rope_sin"] = ("rope", None)
    g["code_predictor"] = c
    v = {
        "rvq.sem.codebook": ("stack", ["code2wav.quantizer.rvq_first.vq.layers.0._codebook.embed"]),
        "rvq.sem.out_proj": ("squeeze_k", "code2wav.quantizer.rvq_first.output_proj.weight"),
        "rvq.aco.codebook": ("stack", [f"code2wav.quantizer.rvq_rest.vq.layers.{j}._codebook.embed" for j in range(spec.cp_steps)]),
        "rvq.aco.out_proj": ("squeeze_k", "code2wav.quantizer.rvq_rest.output_proj.weight"),
        "pre_conv.weight": ("conv", "code2wav.pre_conv.conv.weight"), "pre_conv.bias": ("copy", "code2wav.pre_conv.conv.bias"),
        "pt.norm": ("copy", "code2wav.pre_transformer.norm.weight"),
        "pt.rope_cos": ("rope", None), "pt.rope_sin": ("rope", None),
    }
    for i in range(spec.voc_layers):
        v.update(_layer_rules(f"pt.l{i}", f"code2wav.pre_transformer.layers.{i}", False, True))
    for u, f in enumerate(spec.voc_upsampling_ratios):
        s = f"code2wav.upsample.{u}"
        v.update({
            f"up{u}.tconv.weight": ("tconv", f"{s}.0.conv.weight", f), f"up{u}.tconv.bias": ("copy", f"{s}.0.conv.bias"),
            f"up{u}.dw.weight": ("dw", f"{s}.1.dwconv.conv.weight"), f"up{u}.dw.bias": ("copy", f"{s}.1.dwconv.conv.bias"),
            f"up{u}.ln.weight": ("copy", f"{s}.1.norm.weight"), f"up{u}.ln.bias": ("copy", f"{s}.1.norm.bias"),
            f"up{u}.pw1.weight": ("copy", f"{s}.1.pwconv1.weight"), f"up{u}.pw1.bias": ("copy", f"{s}.1.pwconv1.bias"),
            f"up{u}.pw2.weight": ("copy", f"{s}.1.pwconv2.weight"), f"up{u}.pw2.bias": ("copy", f"{s}.1.pwconv2.bias"),
            f"up{u}.gamma": ("copy", f"{s}.1.gamma"),
        })
    v["dec.conv_in.weight"] = ("conv", "code2wav.decoder.0.conv.weight"); v["dec.conv_in.bias"] = ("copy", "code2wav.decoder.0.conv.bias")
    nb = len(spec.voc_upsample_rates)
    for b, st in enumerate(spec.voc_upsample_rates):
        d = f"code2wav.decoder.{b + 1}.block"
        v.update({
            f"dec.b{b}.snake.alpha": ("copy", f"{d}.0.alpha"), f"dec.b{b}.snake.beta": ("copy", f"{d}.0.beta"),
            f"dec.b{b}.tconv.weight": ("tconv", f"{d}.1.conv.weight", st), f"dec.b{b}.tconv.bias": ("copy", f"{d}.1.conv.bias"),
        })
        for r in range(3):
            q = f"{d}.{r + 2}"
            v.update({
                f"dec.b{b}.r{r}.snake1.alpha": ("copy", f"{q}.act1.alpha"), f"dec.b{b}.r{r}.snake1.beta": ("copy", f"{q}.act1.beta"),
                f"dec.b{b}.r{r}.conv1.weight": ("conv", f"{q}.conv1.conv.weight"), f"dec.b{b}.r{r}.conv1.bias": ("copy", f"{q}.conv1.conv.bias"),
                f"dec.b{b}.r{r}.snake2.alpha": ("copy", f"{q}.act2.alpha"), f"dec.b{b}.r{r}.snake2.beta": ("copy", f"{q}.act2.beta"),
                f"dec.b{b}.r{r}.conv2.weight": ("conv", f"{q}.conv2.conv.weight"), f"dec.b{b}.r{r}.conv2.bias": ("copy", f"{q}.conv2.conv.bias"),
            })
    v["dec.snake_out.alpha"] = ("copy", f"code2wav.decoder.{nb + 1}.alpha"); v["dec.snake_out.beta"] = ("copy", f"code2wav.decoder.{nb + 1}.beta")
    v["dec.conv_out.weight"] = ("conv", f"code2wav.decoder.{nb + 2}.conv.weight"); v["dec.conv_out.bias"] = ("copy", f"code2wav.decoder.{nb + 2}.conv.bias")
    g["tokenizer12hz_decode"] = v
    sp = {"in_conv.weight": ("conv", "speaker_encoder.in_conv.weight"), "in_conv.bias": ("copy", "speaker_encoder.in_conv.bias"),
          "fc.weight": ("copy", "speaker_encoder.fc.weight"), "fc.bias": ("copy", "speaker_encoder.fc.bias")}
    for i in range(spec.spk_layers):
        sp[f"l{i}.conv.weight"] = ("conv", f"speaker_encoder.layers.{i}.conv.weight")
        sp[f"l{i}.conv.bias"] = ("copy", f"speaker_encoder.layers.{i}.conv.bias")
    g["speaker_encoder"] = sp
    return g


def _to_f32(a: np.ndarray) -> np.ndarray:
    if a.dtype == np.uint16:                      # bf16 bits
        return ms.bf16_bits_to_f32(a)
    return np.asarray(a, dtype=np.float32)


def apply_rule(rule, src: dict, missing: list):
    kind = rule[0]
    names = rule[1] if isinstance(rule[1], list) else [rule[1]]
    for n in names:
        if n is not None and n not in src:
            missing.append(n)
    if any(n is not None and n not in src for n in names):
        return None
    if kind == "copy":
        return _to_f32(src[names[0]])
    if kind == "cat0":
        return np.concatenate([_to_f32(src[n]) for n in names], 0)
    if kind == "stack":
        return np.stack([_to_f32(src[n]) for n in names], 0)
    if kind == "conv":
        return conv_fwd(_to_f32(src[names[0]]))
    if kind == "squeeze_k":                       # 1x1 Conv1d [Cout, Cin, 1] -> Linear [Cout, Cin]
        w = _to_f32(src[names[0]])
        return w[:, :, 0] if w.ndim == 3 else w
    if kind == "tconv":
        return tconv_fwd(_to_f32(src[names[0]]), rule[2])
    if kind == "dw":
        return dw_fwd(_to_f32(src[names[0]]))
    raise ValueError(kind)


def convert(src: dict, spec: ms.ModelSpec, out_dir: str, with_speaker: bool = True, overrides: dict | None = None) -> list:
    """-> list of problems (empty = written). Nothing is written unless every tensor checks out."""
    defs = ms.graph_tensor_defs(spec)
    rules = name_map(spec)
    for k, v in (overrides or {}).items():        # "graph/tensor": "source.name" or ["a", "b", ...]
        gname, tname = k.split("/", 1)
        old = rules[gname][tname]
        rules[gname][tname] = (old[0], v) + tuple(old[2:])
    problems, built = [], {}
    graphs = [g for g in ms.GRAPH_FILES if g != "talker_decode"] + (["speaker_encoder"] if with_speaker else [])
    for gname in graphs:
        tensors = []
        for td in defs[gname]:
            rule = rules[gname].get(td.name)
            if rule is None:
                problems.append(f"{gname}/{td.name}: no conversion rule")
                continue
            if rule[0] == "rope":
                arr = ms._gen_tensor(spec, gname, td)                     # tables, not parameters
            else:
                missing: list = []
                arr = apply_rule(rule, src, missing)
                if arr is None:
                    problems.append(f"{gname}/{td.name}: source tensor(s) missing: {missing}")
                    continue
                if tuple(arr.shape) != tuple(td.shape):
                    problems.append(f"{gname}/{td.name}: shape {tuple(arr.shape)} from {rule[1]} != spec {tuple(td.shape)}")
                    continue
                arr = ms.f32_to_bf16_bits(arr).reshape(td.shape) if td.dtype == ms.DT_BF16 else np.ascontiguousarray(arr, np.float32)
            tensors.append((td.name, td.dtype, arr))
        built[gname] = tensors
    if problems:
        return problems
    os.makedirs(out_dir, exist_ok=True)
    for gname, tensors in built.items():
        meta = dict(spec.to_meta()); meta["graph"] = gname; meta["source"] = "converted checkpoint"
        ms.write_lqw(os.path.join(out_dir, gname + ".lqw"), tensors, meta)
    meta = dict(spec.to_meta()); meta["graph"] = "talker_decode"; meta["shares"] = "talker_prefill"
    ms.write_lqw(os.path.join(out_dir, "talker_decode.lqw"), [], meta)
    return []


def export_hf(model_dir: str) -> dict:
    """inverse of convert(): a model directory -> {checkpoint name: fp32 array} (round-trip test / documentation of the map)"""
    spec, graphs = ms.load_model_dir(model_dir)
    rules = name_map(spec)
    out = {}
    for gname, trules in rules.items():
        if gname not in graphs:
            continue
        for tname, rule in trules.items():
            if rule[0] == "rope" or tname not in graphs[gname]:
                continue
            a = _to_f32(np.asarray(graphs[gname][tname]))
            kind, names = rule[0], (rule[1] if isinstance(rule[1], list) else [rule[1]])
            if kind == "copy":
                out[names[0]] = a
            elif kind == "cat0":
                if gname == "talker_prefill":
                    qd, kvd = spec.q_dim, spec.kv_dim
                elif gname == "code_predictor":
                    qd, kvd = spec.cp_q_dim, spec.cp_kv_dim
                else:
                    qd = kvd = spec.voc_heads * spec.voc_head_dim
                out[names[0]], out[names[1]], out[names[2]] = a[:qd], a[qd:qd + kvd], a[qd + kvd:]
            elif kind == "stack":
                for j, n in enumerate(names):
                    out[n] = a[j]
            elif kind == "conv":
                out[names[0]] = conv_inv(a)
            elif kind == "squeeze_k":
                out[names[0]] = a[:, :, None]
            elif kind == "tconv":
                out[names[0]] = tconv_inv(a)
            elif kind == "dw":
                out[names[0]] = dw_inv(a)
    return out


def load_safetensors_dir(d: str) -> dict:
    from safetensors.numpy import load_file
    src = {}
    for fn in sorted(os.listdir(d)):
        if fn.endswith(".safetensors"):
            src.update(load_file(os.path.join(d, fn)))
    return src


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--src", help="directory with *.safetensors")
    ap.add_argument("--src-onnx", help="directory with the reference's seven .onnx files (needs the `onnx` package)")
    ap.add_argument("--out", help="output model directory (.lqw files)")
    ap.add_argument("--spec", default="0.6b", choices=["0.6b", "1.7b", "tiny"])
    ap.add_argument("--map", help="JSON {\"graph/tensor\": \"source name\"} overriding NAME_MAP entries")
    ap.add_argument("--no-speaker-encoder", action="store_true")
    ap.add_argument("--export-hf", help="inverse direction: write DIR/model.safetensors from --from-lqw")
    ap.add_argument("--from-lqw")
    a = ap.parse_args()
    if a.export_hf:
        from safetensors.numpy import save_file
        os.makedirs(a.export_hf, exist_ok=True)
        save_file(export_hf(a.from_lqw), os.path.join(a.export_hf, "model.safetensors"))
        return 0
    if a.src_onnx:
        try:
            import onnx  # noqa: F401
        except ImportError:
            print("--src-onnx needs the `onnx` package (not installed here); convert from the HF safetensors instead", file=sys.stderr)
            return 2
        print("ONNX initialiser import is not implemented: the exported graphs rename and fuse parameters; use --src", file=sys.stderr)
        return 2
    spec = {"0.6b": ms.spec_0p6b, "1.7b": ms.spec_1p7b, "tiny": ms.spec_tiny}[a.spec](0)
    src = load_safetensors_dir(a.src)
    overrides = json.load(open(a.map)) if a.map else None
    problems = convert(src, spec, a.out, with_speaker=not a.no_speaker_encoder, overrides=overrides)
    if problems:
        print(f"{len(problems)} problem(s); nothing written:", file=sys.stderr)
        for p in problems:
            print("  " + p, file=sys.stderr)
        return 1
    print(f"wrote {a.out}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
