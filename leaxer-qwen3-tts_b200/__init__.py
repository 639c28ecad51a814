"""B200-native Qwen3-TTS hot path (drop-in for the seven ONNX Runtime sessions behind
leaxer_qwen::TTSEngine, /root/reference/src/tts_onnx.cpp:545-950).

This package holds only what the path needs:
  csrc/        hand-written sm_100a CUDA kernels + the C-ABI (include/lqt_b200.h)
  host/        C++17 host mirror of the reference interface (tts_onnx.h, CLI, tokenizer, WAV, mel)
  modelspec.py frozen model spec, .lqw weight files, seeded random-init generator
  engine.py    ctypes binding over the C-ABI (what tests/bench call); raises if the CUDA library
               is missing -- there is NO CPU fallback.

The directory name contains '-' (it mirrors the reference repo name), so import it through
`__graft_entry__.load_package()` which registers it as `leaxer_qwen3_tts_b200`.
"""

__all__ = ["modelspec", "engine"]
