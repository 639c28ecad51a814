// WAV reader + linear resampler of the reference host (leaxer-ai/leaxer-qwen3-tts src/io/wav_reader.h:13-21):
// same namespace, names and behaviour; implementation written from that behavioural spec.
#ifndef LEAXER_QWEN_IO_WAV_READER_H
#define LEAXER_QWEN_IO_WAV_READER_H

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace leaxer_qwen {
namespace io {

// mono float32 samples in [-1, 1); empty on any failure; out_sample_rate is written only on success
std::vector<float> read_wav(const std::string& path, int& out_sample_rate);

// linear interpolation; identity when the rates are equal or the input is empty
std::vector<float> resample(const std::vector<float>& audio, int src_sr, int dst_sr);

// 16-bit PCM mono writer of the library (src/io/wav_writer.cpp:31-82, same exported signature): scales by 0.95/peak
// when peak > 1e-4, then clamps; 0 on success, -1 when the file cannot be opened
int write_wav(const char* path, const float* audio, size_t n_samples, int sample_rate);
// the writer the reference CLI uses (src/main_onnx.cpp:15-58): clamp to [-1, 1], int16(sample * 32767.0f), no normalisation
int write_wav_cli(const std::string& path, const std::vector<float>& audio, int sample_rate);

} // namespace io
} // namespace leaxer_qwen

#endif // LEAXER_QWEN_IO_WAV_READER_H
