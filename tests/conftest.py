"""Test plumbing. `-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI exports (no compute).
`-m gpu`: parity tests proper, all through the C-ABI (ctypes over liblqt_b200.so)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from __graft_entry__ import load_package  # noqa: E402

load_package()
from leaxer_qwen3_tts_b200 import modelspec as ms  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def tiny_dir():
    spec = ms.spec_tiny(0)
    return ms.generate_model_dir(ms.default_model_dir(spec), spec)


@pytest.fixture(scope="session")
def full_dir():
    spec = ms.spec_0p6b(0)
    return ms.generate_model_dir(ms.default_model_dir(spec), spec)


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import qwen3_tts_oracle as orc
    return orc


@pytest.fixture(scope="session")
def tiny_oracle(tiny_dir, oracle_mod):
    return oracle_mod.OracleModel(tiny_dir)


@pytest.fixture(scope="session")
def full_oracle(full_dir, oracle_mod):
    return oracle_mod.OracleModel(full_dir)


@pytest.fixture(scope="session")
def full_f32_oracle(full_dir, oracle_mod):
    """oracle without the bf16 KV rounding point (mirrors LQT_KV_F32 parity mode)"""
    return oracle_mod.OracleModel(full_dir, kv_bf16=False)


def _engine(model_dir, **kw):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from leaxer_qwen3_tts_b200 import engine
    return engine.Engine(model_dir, device=0, **kw)


@pytest.fixture(scope="session")
def full_f32_engine(full_dir):
    e = _engine(full_dir, kv_dtype="f32")
    yield e
    e.close()


@pytest.fixture(scope="session")
def tiny_engine(tiny_dir):
    e = _engine(tiny_dir)
    yield e
    e.close()


@pytest.fixture(scope="session")
def full_engine(full_dir):
    e = _engine(full_dir)
    yield e
    e.close()
