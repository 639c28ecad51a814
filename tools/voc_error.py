"""Vocoder waveform error against the CPU oracle (rel-L2) on the 0.6B spec; LQT_CONV_FP32=1 selects the CUDA-core conv kernel.
Test infrastructure: imports oracle/ as the checker. Usage (GPU box): python tools/voc_error.py"""
import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from __graft_entry__ import load_package
load_package()
import numpy as np
from leaxer_qwen3_tts_b200 import engine, modelspec as ms
import importlib.util
spec_o = importlib.util.spec_from_file_location("orc", "/root/repo/oracle/qwen3_tts_oracle.py"); orc = importlib.util.module_from_spec(spec_o); sys.modules["orc"] = orc; spec_o.loader.exec_module(orc)
spec = ms.spec_0p6b(0)
mdir = ms.generate_model_dir(ms.default_model_dir(spec), spec)
m = orc.OracleModel(mdir)
eng = engine.Engine(mdir)
for T in (3, 40):
    codes = np.random.default_rng(T).integers(0, 2048, size=(T, 16))
    ref, n = m.vocoder(codes); ref = ref.numpy()
    out = eng.vocoder_decode(codes)
    print("T", T, "rel_l2 vs oracle", float(np.linalg.norm(out - ref) / np.linalg.norm(ref)))
eng.close()
