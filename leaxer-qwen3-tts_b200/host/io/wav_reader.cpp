// RIFF/WAVE reader and linear resampler (behaviour of src/io/wav_reader.cpp:29-164 of the reference):
//  * PCM (tag 1: 8/16/24/32 bit) and IEEE float (tag 3: 32 bit) only; WAVE_FORMAT_EXTENSIBLE is rejected
//  * chunks are walked in file order, "fmt " beyond 16 bytes is skipped, unknown chunks are skipped without
//    odd-size padding, a short "data" chunk leaves the missing samples at zero
//  * channels are averaged; 64-bit float data decodes to silence
#include "wav_reader.h"

#include <cstdio>
#include <cstring>

namespace leaxer_qwen {
namespace io {

namespace {

struct File {
    std::FILE* f;
    explicit File(const char* p) : f(std::fopen(p, "rb")) {}
    ~File() { if (f) std::fclose(f); }
};

bool get(std::FILE* f, void* dst, size_t n) { return std::fread(dst, 1, n, f) == n; }

uint32_t le32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t le16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

float decode(const unsigned char* p, int bits, bool is_float) {
    switch (bits) {
        case 8:  return ((float)p[0] - 128.0f) / 128.0f;
        case 16: return (float)(int16_t)le16(p) / 32768.0f;
        case 24: {
            int32_t v = (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16));
            if (v & 0x800000) v |= ~0xFFFFFF;                       // sign extension
            return (float)v / 8388608.0f;
        }
        case 32:
            if (is_float) { float v; const uint32_t u = le32(p); std::memcpy(&v, &u, 4); return v; }
            return (float)(int32_t)le32(p) / 2147483648.0f;
        default: return 0.0f;                                       // e.g. 64-bit float: silence
    }
}

}  // namespace

std::vector<float> read_wav(const std::string& path, int& out_sample_rate) {
    File in(path.c_str());
    if (!in.f) return {};
    unsigned char hdr[12];
    if (!get(in.f, hdr, 12) || std::memcmp(hdr, "RIFF", 4) != 0 || std::memcmp(hdr + 8, "WAVE", 4) != 0) return {};

    int tag = 0, channels = 0, rate = 0, bits = 0;
    bool have_fmt = false, have_data = false;
    std::vector<unsigned char> data;
    unsigned char ch[8];
    while (get(in.f, ch, 8)) {
        const uint32_t size = le32(ch + 4);
        if (std::memcmp(ch, "fmt ", 4) == 0) {
            unsigned char f[16];
            if (size < 16 || !get(in.f, f, 16)) return {};
            tag = le16(f); channels = le16(f + 2); rate = (int)le32(f + 4); bits = le16(f + 14);
            if (size > 16) std::fseek(in.f, (long)(size - 16), SEEK_CUR);
            have_fmt = true;
        } else if (std::memcmp(ch, "data", 4) == 0) {
            data.assign(size, 0);
            if (size) { const size_t n = std::fread(data.data(), 1, size, in.f); (void)n; }   // short read: rest stays zero
            have_data = true;
            break;
        } else {
            std::fseek(in.f, (long)size, SEEK_CUR);
        }
    }
    if (!have_fmt || !have_data) return {};
    if (tag != 1 && tag != 3) return {};
    if (channels <= 0 || rate <= 0 || bits <= 0) return {};

    const size_t bps = (size_t)bits / 8;
    if (bps == 0) return {};
    const size_t frames = data.size() / (bps * (size_t)channels);
    std::vector<float> out(frames);
    for (size_t i = 0; i < frames; ++i) {
        float acc = 0.0f;
        for (int c = 0; c < channels; ++c) acc += decode(&data[(i * (size_t)channels + (size_t)c) * bps], bits, tag == 3);
        out[i] = acc / (float)channels;
    }
    out_sample_rate = rate;
    return out;
}

std::vector<float> resample(const std::vector<float>& audio, int src_sr, int dst_sr) {
    if (src_sr == dst_sr || audio.empty()) return audio;
    const double ratio = (double)dst_sr / (double)src_sr;          // same operation order as the reference (:150-160): bit-identical
    const size_t out_len = (size_t)((double)audio.size() * ratio);
    std::vector<float> out(out_len);
    for (size_t i = 0; i < out_len; ++i) {
        const double pos = (double)i / ratio;
        const size_t i0 = (size_t)pos;
        const size_t i1 = (i0 + 1 < audio.size()) ? i0 + 1 : audio.size() - 1;     // right neighbour clamped to the last sample
        const double fr = pos - (double)i0;
        out[i] = (float)((double)audio[i0] * (1.0 - fr) + (double)audio[i1] * fr);
    }
    return out;
}

}  // namespace io
}  // namespace leaxer_qwen
