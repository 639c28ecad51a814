// Log-mel extractor with the numerics of the reference (src/io/mel.cpp): symmetric Hann window (:15-18), HTK mel
// scale, un-normalised triangular filters on floor((n_fft + 1) * f / sr) bins (:32-80), frames without centre padding
// (:185-191), zero-padding to the next power of two, radix-2 decimation-in-time FFT in f32 with twiddle
// cos/sin(-2*pi*k/size) evaluated in f32 (:132-158), power spectrum, natural log of (energy + 1e-10) (:223-232).
// Written from that behavioural spec: the triangles are kept as (left, centre, right) bin triples instead of a dense
// filter matrix and the twiddles of each stage are tabulated once (same f32 values the reference recomputes per butterfly).
#include "mel.h"

#include <algorithm>
#include <cmath>

namespace leaxer_qwen {
namespace io {

namespace {
inline float hz_to_mel(float hz) { return 2595.0f * std::log10(1.0f + hz / 700.0f); }
inline float mel_to_hz(float mel) { return 700.0f * (std::pow(10.0f, mel / 2595.0f) - 1.0f); }
}  // namespace

MelExtractor::MelExtractor(const MelConfig& config) : config_(config) {
    const int win = config_.win_size;
    window_.resize(static_cast<size_t>(std::max(win, 0)));
    for (int i = 0; i < win; ++i)
        window_[static_cast<size_t>(i)] = static_cast<float>(0.5f * (1.0f - std::cos(2.0f * M_PI * i / (win - 1))));   // evaluated in double, as the reference's expression is

    // mel points -> Hz -> FFT bins
    const int bins = config_.n_fft / 2 + 1;
    const float lo = hz_to_mel(config_.fmin), hi = hz_to_mel(config_.fmax);
    std::vector<int> edge(static_cast<size_t>(config_.num_mels + 2));
    for (int i = 0; i < config_.num_mels + 2; ++i) {
        const float mel = lo + (hi - lo) * i / (config_.num_mels + 1);
        const float hz = mel_to_hz(mel);
        const int b = static_cast<int>(std::floor((config_.n_fft + 1) * hz / config_.sample_rate));
        edge[static_cast<size_t>(i)] = std::min(b, bins - 1);
    }
    filters_.resize(static_cast<size_t>(config_.num_mels));
    for (int m = 0; m < config_.num_mels; ++m)
        filters_[static_cast<size_t>(m)] = Tri{edge[static_cast<size_t>(m)], edge[static_cast<size_t>(m) + 1], edge[static_cast<size_t>(m) + 2]};

    // FFT length: n_fft rounded up to a power of two; twiddle tables per stage
    n_pad_ = 1; log2n_ = 0;
    while (n_pad_ < config_.n_fft) { n_pad_ *= 2; ++log2n_; }
    tw_re_.reserve(static_cast<size_t>(n_pad_)); tw_im_.reserve(static_cast<size_t>(n_pad_));
    for (int size = 2; size <= n_pad_; size *= 2) {
        const float step = static_cast<float>(-2.0f * M_PI / size);       // double expression rounded once, then f32 cos/sin per k
        for (int k = 0; k < size / 2; ++k) {
            const float ang = step * k;
            tw_re_.push_back(std::cos(ang));
            tw_im_.push_back(std::sin(ang));
        }
    }
}

void MelExtractor::power_spectrum(std::vector<float>& re, std::vector<float>& im) const {
    const int n = n_pad_;
    for (int i = 0; i < n; ++i) {                        // bit-reversal permutation
        int j = 0;
        for (int b = 0; b < log2n_; ++b) if (i & (1 << b)) j |= 1 << (log2n_ - 1 - b);
        if (j > i) { std::swap(re[static_cast<size_t>(i)], re[static_cast<size_t>(j)]); std::swap(im[static_cast<size_t>(i)], im[static_cast<size_t>(j)]); }
    }
    size_t tw0 = 0;
    for (int size = 2; size <= n; size *= 2) {
        const int half = size / 2;
        for (int base = 0; base < n; base += size) {
            for (int k = 0; k < half; ++k) {
                const float wr = tw_re_[tw0 + static_cast<size_t>(k)], wi = tw_im_[tw0 + static_cast<size_t>(k)];
                const size_t e = static_cast<size_t>(base + k), o = e + static_cast<size_t>(half);
                const float tr = wr * re[o] - wi * im[o];
                const float ti = wr * im[o] + wi * re[o];
                re[o] = re[e] - tr; im[o] = im[e] - ti;
                re[e] = re[e] + tr; im[e] = im[e] + ti;
            }
        }
        tw0 += static_cast<size_t>(half);
    }
}

std::vector<float> MelExtractor::extract(const std::vector<float>& audio) {
    if (audio.empty()) return {};
    const int len = static_cast<int>(audio.size());
    const int win = config_.win_size, hop = config_.hop_size, nfft = config_.n_fft;
    num_frames_ = (len < win) ? 1u : static_cast<size_t>((len - win) / hop + 1);
    const int bins = nfft / 2 + 1, kept = n_pad_ / 2 + 1;
    const int nb = std::min(bins, kept);

    std::vector<float> out(static_cast<size_t>(config_.num_mels) * num_frames_, 0.0f);
    std::vector<float> re(static_cast<size_t>(n_pad_)), im(static_cast<size_t>(n_pad_)), pw(static_cast<size_t>(nb));
    for (size_t t = 0; t < num_frames_; ++t) {
        const int start = static_cast<int>(t) * hop;
        std::fill(re.begin(), re.end(), 0.0f);
        std::fill(im.begin(), im.end(), 0.0f);
        for (int i = 0; i < win && i < nfft; ++i) {
            const int idx = start + i;
            re[static_cast<size_t>(i)] = (idx < len) ? audio[static_cast<size_t>(idx)] * window_[static_cast<size_t>(i)] : 0.0f;
        }
        power_spectrum(re, im);
        for (int k = 0; k < nb; ++k) pw[static_cast<size_t>(k)] = re[static_cast<size_t>(k)] * re[static_cast<size_t>(k)] + im[static_cast<size_t>(k)] * im[static_cast<size_t>(k)];
        for (int m = 0; m < config_.num_mels; ++m) {
            const Tri& f = filters_[static_cast<size_t>(m)];
            // the reference multiplies EVERY bin by its (mostly zero) weight and adds in bin order; zero terms do not
            // change an f32 sum that starts at +0, so only the support of the triangle is visited, in the same order
            float e = 0.0f;
            for (int k = f.left; k < f.center && k < nb; ++k) e += (static_cast<float>(k - f.left) / (f.center - f.left)) * pw[static_cast<size_t>(k)];
            for (int k = f.center; k < f.right && k < nb; ++k) e += (static_cast<float>(f.right - k) / (f.right - f.center)) * pw[static_cast<size_t>(k)];
            out[static_cast<size_t>(m) * num_frames_ + t] = std::log(e + 1e-10f);
        }
    }
    return out;
}

}  // namespace io
}  // namespace leaxer_qwen
