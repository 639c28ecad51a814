// Log-mel front end of the voice-clone path (reference API: leaxer-ai/leaxer-qwen3-tts src/io/mel.h:12-51; behaviour:
// SURVEY.md Appendix D). Same namespace, MelConfig fields/defaults and MelExtractor public interface.
#ifndef LEAXER_QWEN_IO_MEL_H
#define LEAXER_QWEN_IO_MEL_H

#include <cstddef>
#include <vector>

namespace leaxer_qwen {
namespace io {

struct MelConfig {
    int sample_rate = 24000;
    int n_fft = 1024;
    int hop_size = 256;
    int win_size = 1024;
    int num_mels = 128;
    float fmin = 0.0f;
    float fmax = 12000.0f;
};

class MelExtractor {
public:
    explicit MelExtractor(const MelConfig& config);

    // log-mel spectrogram, row-major [num_mels][num_frames]; no centre padding; natural log of (energy + 1e-10)
    std::vector<float> extract(const std::vector<float>& audio);

    size_t num_frames() const { return num_frames_; }
    size_t num_mels() const { return static_cast<size_t>(config_.num_mels); }

private:
    struct Tri { int left, center, right; };      // triangular filter on FFT bins: up [left, center), down [center, right)
    MelConfig config_;
    std::vector<float> window_;                   // symmetric Hann, denominator win - 1
    std::vector<Tri> filters_;
    std::vector<float> tw_re_, tw_im_;            // twiddles of every butterfly stage, concatenated (size n - 1)
    int n_pad_ = 0, log2n_ = 0;
    size_t num_frames_ = 0;

    void power_spectrum(std::vector<float>& re, std::vector<float>& im) const;   // in-place radix-2 FFT of re (im = 0 on entry)
};

} // namespace io
} // namespace leaxer_qwen

#endif // LEAXER_QWEN_IO_MEL_H
