// TEST INFRASTRUCTURE (oracle side): a tiny driver around the io/ API shared by the reference and this repo
// (leaxer_qwen::io: MelExtractor, read_wav, resample, tokenizer). It is compiled twice:
//   oracle/_ref/io_dump_ref  <- against the reference's own sources where they lie (/root/reference/src/io/*.cpp)
//   host/build/io_dump       <- against this repo's re-implementation (leaxer-qwen3-tts_b200/host/io/*.cpp)
// and tests/ compare the two byte streams (and the committed fixtures made from the reference build).
// Output: raw little-endian values on stdout. Commands:
//   mel N SEED            log-mel [128][frames] of N synthetic samples (24 kHz, n_fft 1024, hop 256: src/tts_onnx.cpp:347-354)
//   wav PATH              int32 sample rate (or -1) followed by the decoded samples
//   resample N SRC DST    resample(N synthetic samples)
//   tok VOCAB MERGES TEXT int32 token ids ("-" for a missing file: tokenizer without vocab/merges)
//   tokhf VOCAB MERGES FILE   (this repo's build only) HF-faithful tokenizer mode: ids of every line of FILE
//   melwav PATH           the clone front-end of src/tts_onnx.cpp:331-359: read_wav -> resample to 24 kHz -> log-mel [128][frames]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mel.h"
#include "tokenizer.h"
#include "wav_reader.h"

using namespace leaxer_qwen::io;

static std::vector<float> synth(int n, unsigned seed) {
    std::vector<float> a(static_cast<size_t>(n));
    uint32_t s = seed * 2654435761u + 12345u;
    for (int i = 0; i < n; ++i) {
        s = s * 1664525u + 1013904223u;
        const float noise = (static_cast<float>(s >> 8) / 16777216.0f - 0.5f) * 0.1f;
        const double t = static_cast<double>(i) / 24000.0;
        a[static_cast<size_t>(i)] = static_cast<float>(0.4 * std::sin(2 * M_PI * 220.0 * t) + 0.25 * std::sin(2 * M_PI * 1330.0 * t) + 0.1 * std::sin(2 * M_PI * 5100.0 * t)) + noise;
    }
    return a;
}
static void put_f(const std::vector<float>& v) { if (!v.empty()) std::fwrite(v.data(), 4, v.size(), stdout); }
static void put_i(int32_t v) { std::fwrite(&v, 4, 1, stdout); }

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string cmd = argv[1];
    if (cmd == "mel" && argc >= 4) {
        MelConfig mc;
        mc.sample_rate = 24000; mc.n_fft = 1024; mc.hop_size = 256; mc.win_size = 1024; mc.num_mels = 128; mc.fmin = 0.0f; mc.fmax = 12000.0f;
        MelExtractor mel(mc);
        const std::vector<float> m = mel.extract(synth(std::atoi(argv[2]), static_cast<unsigned>(std::atoi(argv[3]))));
        put_i(static_cast<int32_t>(mel.num_frames()));
        put_f(m);
        return 0;
    }
    if (cmd == "melwav" && argc >= 3) {
        int sr = -1;
        std::vector<float> a = read_wav(argv[2], sr);
        if (a.empty()) { put_i(-1); return 0; }
        if (sr != 24000) a = resample(a, sr, 24000);
        MelConfig mc;
        mc.sample_rate = 24000; mc.n_fft = 1024; mc.hop_size = 256; mc.win_size = 1024; mc.num_mels = 128; mc.fmin = 0.0f; mc.fmax = 12000.0f;
        MelExtractor mel(mc);
        const std::vector<float> m = mel.extract(a);
        put_i(static_cast<int32_t>(mel.num_frames()));
        put_f(m);
        return 0;
    }
    if (cmd == "wav" && argc >= 3) {
        int sr = -1;
        const std::vector<float> a = read_wav(argv[2], sr);
        put_i(sr);
        put_f(a);
        return 0;
    }
    if (cmd == "resample" && argc >= 5) {
        put_f(resample(synth(std::atoi(argv[2]), 7u), std::atoi(argv[3]), std::atoi(argv[4])));
        return 0;
    }
    if (cmd == "tok" && argc >= 5) {
        if (std::strcmp(argv[2], "-") != 0) load_vocab(argv[2]);
        if (std::strcmp(argv[3], "-") != 0) load_merges(argv[3]);
        put_i(is_tokenizer_ready() ? 1 : 0);
        for (int32_t t : tokenize(argv[4])) put_i(t);
        return 0;
    }
#ifdef LEAXER_HAS_HF_TOKENIZER
    // tokhf VOCAB MERGES FILE: one text per line of FILE ("\\n" / "\\r" / "\\t" / "\\\\" escapes) through the HF-faithful mode (this repo's
    // host only); output per line: count, ids
    if (cmd == "tokhf" && argc >= 5) {
        set_tokenizer_mode(TokenizerMode::HF);
        if (!load_vocab(argv[2]) || !load_merges(argv[3])) return 1;
        FILE* f = std::fopen(argv[4], "rb");
        if (!f) return 1;
        std::string all;
        char buf[4096];
        size_t got;
        while ((got = std::fread(buf, 1, sizeof(buf), f)) > 0) all.append(buf, got);
        std::fclose(f);
        size_t pos = 0;
        while (pos <= all.size()) {
            size_t e = all.find('\n', pos);
            if (e == std::string::npos) { if (pos == all.size()) break; e = all.size(); }
            std::string text;
            for (size_t i = pos; i < e; ++i) {
                if (all[i] == '\\' && i + 1 < e) {
                    const char c = all[++i];
                    text.push_back(c == 'n' ? '\n' : c == 'r' ? '\r' : c == 't' ? '\t' : c);
                } else text.push_back(all[i]);
            }
            const std::vector<int32_t> ids = tokenize(text);
            put_i(static_cast<int32_t>(ids.size()));
            for (int32_t t : ids) put_i(t);
            pos = e + 1;
        }
        return 0;
    }
#endif
#ifdef LEAXER_HAS_NFC
    // nfc FILE: every line of FILE (same escapes as tokhf) through normalize_nfc; output: the normalised lines, '\n'-terminated (a '\n'
    // inside a text is written as the two characters backslash n again)
    if (cmd == "nfc" && argc >= 3) {
        FILE* f = std::fopen(argv[2], "rb");
        if (!f) return 1;
        std::string all;
        char buf[4096];
        size_t got;
        while ((got = std::fread(buf, 1, sizeof(buf), f)) > 0) all.append(buf, got);
        std::fclose(f);
        size_t pos = 0;
        while (pos < all.size()) {
            size_t e = all.find('\n', pos);
            if (e == std::string::npos) e = all.size();
            std::string text;
            for (size_t i = pos; i < e; ++i) {
                if (all[i] == '\\' && i + 1 < e) {
                    const char c = all[++i];
                    text.push_back(c == 'n' ? '\n' : c == 'r' ? '\r' : c == 't' ? '\t' : c);
                } else text.push_back(all[i]);
            }
            for (char c : normalize_nfc(text)) {
                if (c == '\n') std::fputs("\\n", stdout);
                else if (c == '\\') std::fputs("\\\\", stdout);
                else std::fputc(c, stdout);
            }
            std::fputc('\n', stdout);
            pos = e + 1;
        }
        return 0;
    }
#endif
    return 2;
}
