#!/usr/bin/env python
"""Bisecting aid for the persistent frame kernel: 1-layer talker, one prefill row; compares the exchange buffers
(qkv, x1, act, x) left by the kernel with the oracle's intermediate values. GPU box only.
Usage: python tools/fk_debug_layer.py [tiny|0.6b]"""
import dataclasses
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

load_package()
from leaxer_qwen3_tts_b200 import engine, modelspec as ms  # noqa: E402
from oracle import qwen3_tts_oracle as orc  # noqa: E402

base = {"tiny": ms.spec_tiny, "0.6b": ms.spec_0p6b}[sys.argv[1] if len(sys.argv) > 1 else "0.6b"](0)
spec = dataclasses.replace(base, name=base.name + "-1layer", layers=1)
mdir = ms.generate_model_dir(ms.default_model_dir(spec), spec)
m = orc.OracleModel(mdir)
eng = engine.Engine(mdir)
H, D = spec.hidden, spec.head_dim
x = np.random.default_rng(1).standard_normal((1, H)).astype(np.float32)
logits, hid = eng.talker_prefill(x, slot=0)
g = m.g["talker_prefill"]
xt = torch.from_numpy(x)
h = orc.rmsnorm(xt, g["l0.ln1"], spec.rms_eps)
qkv = h @ g["l0.wqkv"].T
qd, kvd = spec.heads * D, spec.kv_heads * D
v = qkv[:, qd + kvd:].reshape(1, spec.kv_heads, D)
if m.kv_bf16:
    v = orc.bf16_round(v)
o = v.repeat_interleave(spec.heads // spec.kv_heads, dim=1).reshape(1, qd)     # one position: attention output = v
x1 = xt + o @ g["l0.wo"].T
h2 = orc.rmsnorm(x1, g["l0.ln2"], spec.rms_eps)
act = F.silu(h2 @ g["l0.wgate"].T) * (h2 @ g["l0.wup"].T)
xo = x1 + act @ g["l0.wdown"].T
hd = orc.rmsnorm(xo, g["norm"], spec.rms_eps)
lg = hd @ g["head"].T


def cmp(name, got, ref):
    ref = ref.numpy().reshape(-1)
    e = np.abs(got - ref)
    bad = np.nonzero(e > 1e-3 * (1 + np.abs(ref)))[0]
    print(f"{name:8s} n={len(ref):5d} max|err|={e.max():.3e} at {int(e.argmax())}  bad={len(bad)}  first bad={bad[:12].tolist()}  got/ref at worst: {got[e.argmax()]:.5f} / {ref[e.argmax()]:.5f}")


cmp("qkv", eng.debug_exchange(1, qd + 2 * kvd), qkv)
cmp("x1", eng.debug_exchange(2, H), x1)
cmp("act", eng.debug_exchange(3, spec.inter), act)
cmp("x", eng.debug_exchange(0, H), xo)
cmp("hidden", hid, hd)
cmp("logits", logits, lg)
eng.close()
