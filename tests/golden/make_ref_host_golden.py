#!/usr/bin/env python
"""Writes tests/golden/ref_host_golden.npz from the REFERENCE's own host code: oracle/_ref/tts_host_ref =
/root/reference/src/tts_onnx.cpp compiled unmodified against the ORT shim (oracle/Makefile). Run in the container that has
/root/reference mounted:  make -C oracle && python tests/golden/make_ref_host_golden.py
Stored per case: sha256 of the per-call trace and the RESULT line; the reference's filter outputs; the reference's log-mel
of the synthetic clone clip."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_host_cases as rc  # noqa: E402


def main():
    out = {}
    with tempfile.TemporaryDirectory() as d:
        mdir = rc.make_model_dir(d)
        for name, (lang, k, mx, eos, text) in rc.ID_CASES.items():
            tr, res = rc.run_ref(["ids", mdir, lang, 0.8, k, 0.95, mx, eos, *rc.wrap(text)])
            out[f"{name}_sha"], out[f"{name}_result"] = np.asarray(rc.sha(tr)), np.asarray(res)
        for name, (lang, k, mx, eos, text) in rc.TEXT_CASES.items():
            tr, res = rc.run_ref(["text", mdir, lang, 0.8, k, 0.95, mx, eos, text])
            out[f"{name}_sha"], out[f"{name}_result"] = np.asarray(rc.sha(tr)), np.asarray(res)
        wav = rc.write_ref_wav(os.path.join(d, "ref3s.wav"))
        for name, (lang, k, mx, eos, text) in rc.CLONE_CASES.items():
            tr, res = rc.run_ref(["clone", mdir, lang, 0.8, k, 0.95, mx, eos, wav, text])
            out[f"{name}_sha"], out[f"{name}_result"] = np.asarray(rc.sha(tr)), np.asarray(res)
        out["clone_mel"] = rc.ref_melwav(rc.IO_REF, wav)
        fin, fout = os.path.join(d, "in.f32"), os.path.join(d, "out.f32")
        for V, k, p, seed in rc.FILTER_CASES:
            x = rc.filter_logits(V, seed, ties=(seed % 3 == 2))
            x.tofile(fin)
            subprocess.run([rc.HOST_REF, "filt", fin, fout, str(k), str(p)], check=True)
            a = np.fromfile(fout, np.float32)[:V]
            a.tofile(fin)
            subprocess.run([rc.HOST_REF, "filt", fin, fout, str(k), str(p)], check=True)
            o = np.fromfile(fout, np.float32)
            out[f"filt_{V}_{k}_{seed}"] = np.stack([a, o[V:2 * V], o[2 * V:]])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_host_golden.npz"), **out)
    print(f"wrote {len(out)} entries from {rc.HOST_REF}")


if __name__ == "__main__":
    main()
