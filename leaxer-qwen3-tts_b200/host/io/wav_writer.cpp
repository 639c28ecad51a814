// 16-bit PCM mono WAV writers.
//  * io::write_wav: the library writer of the reference (src/io/wav_writer.cpp:31-82): scales by 0.95/peak when peak > 1e-4
//  * write_wav_cli: the writer the reference CLI actually uses (src/main_onnx.cpp:15-58): clamps to [-1, 1],
//    int16(sample * 32767.0f) with truncation toward zero, no normalisation; returns -1 when the file cannot be opened.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "wav_reader.h"

namespace leaxer_qwen {
namespace io {

namespace {
void put32(unsigned char* p, uint32_t v) { p[0] = v & 255; p[1] = (v >> 8) & 255; p[2] = (v >> 16) & 255; p[3] = (v >> 24) & 255; }
void put16(unsigned char* p, uint16_t v) { p[0] = v & 255; p[1] = (v >> 8) & 255; }

bool write_pcm16(const std::string& path, const std::vector<int16_t>& pcm, int rate) {
    std::FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const uint32_t bytes = (uint32_t)(pcm.size() * 2);
    unsigned char h[44];
    std::copy_n("RIFF", 4, h); put32(h + 4, 36 + bytes); std::copy_n("WAVE", 4, h + 8);
    std::copy_n("fmt ", 4, h + 12); put32(h + 16, 16); put16(h + 20, 1); put16(h + 22, 1);
    put32(h + 24, (uint32_t)rate); put32(h + 28, (uint32_t)rate * 2); put16(h + 32, 2); put16(h + 34, 16);
    std::copy_n("data", 4, h + 36); put32(h + 40, bytes);
    bool ok = std::fwrite(h, 1, 44, f) == 44;
    std::vector<unsigned char> b(pcm.size() * 2);
    for (size_t i = 0; i < pcm.size(); ++i) put16(&b[2 * i], (uint16_t)pcm[i]);
    if (!b.empty()) ok = ok && std::fwrite(b.data(), 1, b.size(), f) == b.size();
    std::fclose(f);
    return ok;
}
}  // namespace

int write_wav(const char* path, const float* audio, size_t n_samples, int sample_rate) {
    float peak = 0.0f;
    for (size_t i = 0; i < n_samples; ++i) peak = std::max(peak, std::fabs(audio[i]));
    const float scale = (peak > 1e-4f) ? 0.95f / peak : 1.0f;
    std::vector<int16_t> pcm(n_samples);
    for (size_t i = 0; i < n_samples; ++i) {
        const float v = std::max(-1.0f, std::min(1.0f, audio[i] * scale));
        pcm[i] = (int16_t)(v * 32767.0f);
    }
    return write_pcm16(path ? path : "", pcm, sample_rate) ? 0 : -1;
}

int write_wav_cli(const std::string& path, const std::vector<float>& audio, int sample_rate) {
    std::vector<int16_t> pcm(audio.size());
    for (size_t i = 0; i < audio.size(); ++i) {
        const float v = std::max(-1.0f, std::min(1.0f, audio[i]));
        pcm[i] = (int16_t)(v * 32767.0f);
    }
    return write_pcm16(path, pcm, sample_rate) ? 0 : -1;
}

}  // namespace io
}  // namespace leaxer_qwen
