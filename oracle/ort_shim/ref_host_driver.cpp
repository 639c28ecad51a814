// TEST INFRASTRUCTURE -- drives the REFERENCE's own TTSEngine (src/tts_onnx.cpp, compiled unmodified from /root/reference
// by oracle/Makefile) over the stub graphs of the ORT shim, and prints the per-call trace. tests/test_ref_host_pin.py
// compares that trace, the sampler filters and the drawn tokens with oracle/qwen3_tts_oracle.py run over the same stubs.
//
//   tts_host_ref ids   MODEL_DIR LANG TEMP TOPK TOPP MAXNEW EOS_AT ID...      public synthesize_tokens()
//   tts_host_ref text  MODEL_DIR LANG TEMP TOPK TOPP MAXNEW EOS_AT TEXT       public synthesize() (tokenizer files next to MODEL_DIR)
//   tts_host_ref clone MODEL_DIR LANG TEMP TOPK TOPP MAXNEW EOS_AT WAV TEXT   public synthesize_clone()
//   tts_host_ref filt  IN.f32 OUT.f32 K P      out = [top_k_filter(x) | softmax(x) | top_p_filter(softmax(x))], the reference's statics
//   tts_host_ref draw  MODEL_DIR IN.f32 TEMP TOPK TOPP N                      N x sample_token(x) (std::mt19937: unseeded)
// LANG: auto|en|zh|ja|ko. EOS_AT: attention-mask length at which the decode stub favours CODEC_EOS (-1: never).
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#define ORT_SHIM_IMPLEMENTATION
#include "stub_graphs.h"

#define private public          // the sampler statics and sample_token are private members of the reference class
#include "tts_onnx.h"
#undef private

using namespace leaxer_qwen;

static Language parse_lang(const std::string& s) {
    if (s == "en") return Language::English;
    if (s == "zh") return Language::Chinese;
    if (s == "ja") return Language::Japanese;
    if (s == "ko") return Language::Korean;
    return Language::Auto;
}
static std::vector<float> read_f32(const char* path) {
    std::ifstream f(path, std::ios::binary);
    std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::vector<float> v(raw.size() / 4);
    std::memcpy(v.data(), raw.data(), v.size() * 4);
    return v;
}
static void finish(const std::vector<float>& audio) {
    std::fputs(ort_shim::state().trace.c_str(), stdout);
    const ort_shim::Digest d = ort_shim::digest_words(reinterpret_cast<const uint32_t*>(audio.data()), audio.size(), 0);
    std::printf("RESULT audio %zu %08x %08x calls %ld\n", audio.size(), d.s, d.x, ort_shim::state().calls);
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string cmd = argv[1];
    if (cmd == "filt" && argc >= 6) {
        const std::vector<float> x = read_f32(argv[2]);
        std::vector<float> a = x, b = x;
        TTSEngine::top_k_filter(a, std::atoi(argv[4]));
        TTSEngine::softmax(b);
        std::vector<float> c = b;
        TTSEngine::top_p_filter(c, (float)std::atof(argv[5]));
        std::ofstream o(argv[3], std::ios::binary);
        o.write(reinterpret_cast<const char*>(a.data()), a.size() * 4);
        o.write(reinterpret_cast<const char*>(b.data()), b.size() * 4);
        o.write(reinterpret_cast<const char*>(c.data()), c.size() * 4);
        return 0;
    }
    if (argc < 3) return 2;
    TTSEngine eng(argv[2]);
    if (!eng.is_ready()) { std::printf("NOT_READY %s\n", eng.get_error().c_str()); return 3; }
    if (cmd == "draw" && argc >= 8) {
        const std::vector<float> x = read_f32(argv[3]);
        SamplingParams sp;
        sp.temperature = (float)std::atof(argv[4]); sp.top_k = std::atoi(argv[5]); sp.top_p = (float)std::atof(argv[6]);
        std::printf("TOKENS");
        for (int i = 0; i < std::atoi(argv[7]); ++i) std::printf(" %lld", (long long)eng.sample_token(x, sp));
        std::printf("\n");
        return 0;
    }
    if (argc < 10) return 2;
    const Language lang = parse_lang(argv[3]);
    SamplingParams sp;
    sp.temperature = (float)std::atof(argv[4]); sp.top_k = std::atoi(argv[5]); sp.top_p = (float)std::atof(argv[6]);
    sp.max_new_tokens = std::atoi(argv[7]);
    ort_shim::state().eos_at = std::atol(argv[8]);
    std::printf("READY speaker_encoder %d\n", eng.has_speaker_encoder() ? 1 : 0);
    if (cmd == "ids") {
        std::vector<int64_t> ids;
        for (int i = 9; i < argc; ++i) ids.push_back(std::atoll(argv[i]));
        finish(eng.synthesize_tokens(ids, lang, sp));
    } else if (cmd == "text") {
        finish(eng.synthesize(argv[9], lang, sp));
    } else if (cmd == "clone" && argc >= 11) {
        finish(eng.synthesize_clone(argv[10], argv[9], lang, sp));
    } else {
        return 2;
    }
    return 0;
}
