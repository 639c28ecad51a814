"""io/ utilities of the C++ host (tokenizer, WAV reader, resampler, log-mel) against the reference's own sources.
The reference keeps these components "unchanged" (north_star); here they are re-implemented from the behavioural spec
(SURVEY.md Appendix D), so the checker is the reference code itself: fixtures generated from it
(tests/golden/io_reference.npz, made by tests/golden/make_io_golden.py) and, where /root/reference is mounted, the live
reference build (oracle/_ref/io_dump_ref). Integer results (token ids, sample rates, frame counts) must be identical;
float results bit-identical for the reader/resampler, within 1e-4 absolute for log-mel (libm differences)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import io_cases  # noqa: E402
import make_io_golden  # noqa: E402

HOST = os.path.join(ROOT, "leaxer-qwen3-tts_b200", "host")
MINE = os.path.join(HOST, "build", "io_dump")
REF = make_io_golden.REF


@pytest.fixture(scope="module")
def mine(tmp_path_factory):
    subprocess.run(["make", "-s", "-C", HOST, "build/io_dump"], check=True)
    return make_io_golden.collect(MINE, str(tmp_path_factory.mktemp("io_mine")))


def _check(mine, ref):
    assert set(mine) == set(ref)
    for k in sorted(ref):
        a, b = mine[k], ref[k]
        assert a.shape == b.shape, (k, a.shape, b.shape)
        if a.dtype.kind == "i":
            assert np.array_equal(a, b), k
        elif k.startswith("mel_"):
            assert np.allclose(a, b, atol=1e-4, rtol=0), (k, float(np.abs(a - b).max()))
        else:
            assert np.array_equal(a, b), (k, float(np.abs(a - b).max()) if a.size else 0.0)


def test_io_matches_reference_fixtures(mine):
    ref = dict(np.load(os.path.join(ROOT, "tests", "golden", "io_reference.npz")))
    _check(mine, ref)


@pytest.mark.skipif(not os.path.exists("/root/reference/src/io/mel.cpp"), reason="reference sources not mounted")
def test_io_matches_live_reference_build(mine, tmp_path):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True, stderr=subprocess.DEVNULL)
    _check(mine, make_io_golden.collect(REF, str(tmp_path)))


def test_io_expected_shapes(mine):
    """the reference's own unit-test expectations (tests/test_mel.cpp, tests/test_wav_reader.cpp)"""
    assert int(mine["mel_72000_1_frames"][0]) == 278 and mine["mel_72000_1"].size == 128 * 278      # 3 s reference clip
    assert int(mine["mel_700_4_frames"][0]) == 1                                                      # shorter than the window
    assert int(mine["wav_s16_mono_24k_sr"][0]) == 24000 and mine["wav_s16_mono_24k"].size == 1000
    assert mine["wav_s16_stereo_44k"].size == 1000                                                    # channels averaged
    assert mine["wav_extensible_rejected"].size == 0 and int(mine["wav_extensible_rejected_sr"][0]) == -1
    assert mine["wav_missing_file"].size == 0
    assert np.all(mine["wav_f64_mono_silence"] == 0)
    assert mine["resample_1000_16000_24000"].size == 1500
    assert mine["tok_0"].tolist()[0] == 1 and mine["tok_0"].tolist()[1:] == [1025, 1030]              # hello, Ġworld
    assert mine["tok_novocab_0"].tolist()[1:] == list(b"hello world")                                 # raw bytes without a vocab


def test_cli_argument_handling(tmp_path):
    """src/main_onnx.cpp:89-135: help -> 0; missing -m/-p -> usage + 1; missing model directory -> 1"""
    subprocess.run(["make", "-s", "-C", HOST], check=True)
    cli = os.path.join(HOST, "leaxer-qwen-b200")
    r = subprocess.run([cli, "--help"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "--max-tokens N" in r.stdout and "--ref PATH" in r.stdout
    r = subprocess.run([cli, "-p", "hi"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "--model and --prompt are required" in r.stderr
    r = subprocess.run([cli, "-m", str(tmp_path / "nope"), "-p", "hi", "--lang"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "model directory not found" in r.stderr
    # existing directory without model files: the engine reports the load failure (no GPU needed to get that far)
    r = subprocess.run([cli, "-m", str(tmp_path), "-p", "hi", "-o", str(tmp_path / "sub" / "o.wav")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Error:" in r.stderr and "Model:" in r.stdout and (tmp_path / "sub").is_dir()


def test_file_parsers_survive_fuzzing_under_asan(tmp_path):
    """host/io/wav_reader.cpp (the --ref clip of the clone path) and the tokenizer's vocab.json / merges.txt loaders compiled with
    -fsanitize=address,undefined into tests/native/host_io_fuzz.cpp and fed mutated, truncated and shuffled copies of well-formed files:
    every one is read or rejected; an out-of-bounds access, an overflow or a crash aborts the driver."""
    import shutil
    import subprocess
    import numpy as np
    import io_cases
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    io_dir = os.path.join(ROOT, "leaxer-qwen3-tts_b200", "host", "io")
    exe = str(tmp_path / "host_io_fuzz")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-I" + io_dir, "-o", exe,
                        os.path.join(ROOT, "tests", "native", "host_io_fuzz.cpp"), os.path.join(io_dir, "wav_reader.cpp"), os.path.join(io_dir, "tokenizer.cpp")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 and "sanitize" in r.stdout:
        pytest.skip("sanitizer runtime not available: " + r.stdout[-200:])
    assert r.returncode == 0, r.stdout[-2000:]
    t = np.arange(2400) / 24000.0
    mono, stereo = str(tmp_path / "mono.wav"), str(tmp_path / "stereo.wav")
    io_cases._wav(mono, 1, 1, 24000, 16, ((0.4 * np.sin(2 * np.pi * 220 * t)) * 32767).astype("<i2").tobytes())
    io_cases._wav(stereo, 3, 2, 44100, 32, (0.4 * np.sin(2 * np.pi * 220 * np.arange(4800) / 44100.0)).astype("<f4").tobytes())
    vp, mp = io_cases.write_tokenizer_files(str(tmp_path / "tok"))
    for what, src, iters, seed in (("wav", mono, 5000, 1), ("wav", stereo, 5000, 2), ("vocab", vp, 3000, 3), ("merges", mp, 3000, 4)):
        r = subprocess.run([exe, what, src, str(tmp_path / "scratch.bin"), str(iters), str(seed)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           timeout=600, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
        last = r.stdout.decode("utf-8", "replace").strip().splitlines()[-1:] or [""]
        assert r.returncode == 0 and last[0].startswith("ok "), (what, r.returncode, r.stderr.decode("utf-8", "replace")[-3000:])
        ok, bad = (int(x) for x in last[0].split()[1:3])
        assert ok + bad == iters and ok > 0 and bad > 0, (what, ok, bad)
