// Host side of leaxer_qwen::TTSEngine over the C-ABI (include/lqt_b200.h). Mirrors the control flow and
// the error conventions of the reference (src/tts_onnx.cpp:84-130, 238-436): the constructor never
// throws, failures leave ready_ == false with error_msg_ set; synthesis failures return an empty
// vector and log to stderr with the "[TTSEngine]" prefix.
#include "tts_onnx.h"

#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <filesystem>
#include <iostream>
#include <random>

#include "../../include/lqt_b200.h"
#include "io/mel.h"
#include "io/tokenizer.h"
#include "io/wav_reader.h"

namespace leaxer_qwen {

namespace fs = std::filesystem;

Speaker parse_speaker(const std::string& name) {           // src/tts_onnx.cpp:54-68
    std::string s(name);
    for (char& ch : s) ch = static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
    static const struct { const char* key; Speaker value; } table[] = {
        {"serena", Speaker::Serena}, {"vivian", Speaker::Vivian}, {"uncle_fu", Speaker::Uncle_Fu},
        {"unclefu", Speaker::Uncle_Fu}, {"dylan", Speaker::Dylan}, {"eric", Speaker::Eric},
        {"ryan", Speaker::Ryan}, {"aiden", Speaker::Aiden}, {"ono_anna", Speaker::Ono_Anna},
        {"onoanna", Speaker::Ono_Anna}, {"sohee", Speaker::Sohee}};
    for (const auto& e : table)
        if (s == e.key) return e.value;
    return Speaker::None;
}

TTSEngine::TTSEngine(const std::string& model_dir) : model_dir_(model_dir) {
    try {
        if (const char* env = std::getenv("LEAXER_SEED")) seed_ = static_cast<uint32_t>(std::strtoul(env, nullptr, 10));
        else seed_ = std::random_device{}();

        int device = 0;
        if (const char* env = std::getenv("LEAXER_DEVICE")) device = std::atoi(env);
        if (lqt_create(model_dir.c_str(), device, &handle_) != 0 || !handle_) {
            // reference: "Failed to load required ONNX models" (src/tts_onnx.cpp:100-104)
            error_msg_ = std::string("Failed to load required ONNX models (") + lqt_create_error() + ")";
            handle_ = nullptr;
            return;
        }
        lqt_info info{};
        lqt_get_info(handle_, &info);
        has_speaker_encoder_ = info.has_speaker_encoder != 0;

        // tokenizer files live where the reference looks for them (:110-121); missing files only warn
        const fs::path base = fs::path(model_dir_).parent_path() / "models" / "Qwen3-TTS-12Hz-0.6B-Base";
        const fs::path vocab = base / "vocab.json", merges = base / "merges.txt";
        if (fs::exists(vocab) && fs::exists(merges)) {
            if (!io::load_vocab(vocab.string()) || !io::load_merges(merges.string())) {
                error_msg_ = "Failed to load tokenizer";
                return;
            }
        } else {
            std::cerr << "[TTSEngine] Warning: Tokenizer not found at " << base << std::endl;
        }
        ready_ = true;
    } catch (const std::exception& e) {
        error_msg_ = std::string("Error: ") + e.what();
    }
}

TTSEngine::~TTSEngine() {
    if (handle_) lqt_destroy(handle_);
}

// [IM_START, ASSISTANT, TTS_BOS, ...text..., TTS_EOS, IM_END]  (src/tts_onnx.cpp:243-259)
std::vector<int64_t> TTSEngine::wrap_text(const std::string& text, bool& ok) {
    std::vector<int64_t> ids = {config::IM_START, config::ASSISTANT, config::TTS_BOS};
    ok = io::is_tokenizer_ready();
    if (!ok) {
        std::cerr << "[TTSEngine] Tokenizer not ready" << std::endl;
        return {};
    }
    for (int32_t t : io::tokenize(text)) ids.push_back(static_cast<int64_t>(t));
    ids.push_back(config::TTS_EOS);
    ids.push_back(config::IM_END);
    return ids;
}

std::vector<float> TTSEngine::run_tokens(const std::vector<int64_t>& ids, Language lang, const SamplingParams& params,
                                         const float* speaker_embed, const ChunkCallback* on_chunk) {
    last_codes_.clear();
    if (ids.size() < 5) {       // the reference indexes ids[3] and ids[n-2] unchecked (:493, :516); refuse instead
        std::cerr << "[TTSEngine] Synthesis error: token sequence too short" << std::endl;
        return {};
    }
    lqt_info info{};
    lqt_get_info(handle_, &info);
    lqt_sampling sp{};
    sp.temperature = params.temperature; sp.top_p = params.top_p; sp.top_k = params.top_k;
    sp.max_new_tokens = std::max(0, std::min(params.max_new_tokens, info.max_pos - 16));
    sp.seed = seed_; sp.utterance_id = utterance_++; sp.greedy = 0;
    const size_t cap = static_cast<size_t>(std::max(sp.max_new_tokens, 1));
    std::vector<float> audio(cap * info.samples_per_frame);
    std::vector<int64_t> codes(cap * 16);
    int64_t n_samples = 0;
    int32_t n_frames = 0;
    auto trampoline = [](void* user, const float* pcm, int64_t first, int64_t n) {
        (*static_cast<const ChunkCallback*>(user))(pcm, first, n);
    };
    const int rc = lqt_synthesize_stream(handle_, ids.data(), static_cast<int32_t>(ids.size()),
                                         static_cast<int32_t>(language_to_codec_id(lang)), speaker_embed, &sp,
                                         audio.data(), static_cast<int64_t>(audio.size()), &n_samples,
                                         codes.data(), &n_frames, on_chunk ? +trampoline : nullptr,
                                         const_cast<ChunkCallback*>(on_chunk));
    if (rc != 0) {
        std::cerr << "[TTSEngine] Synthesis error: " << lqt_last_error(handle_) << std::endl;
        return {};
    }
    codes.resize(static_cast<size_t>(n_frames) * 16);
    last_codes_ = std::move(codes);
    audio.resize(static_cast<size_t>(n_samples));       // empty when the first code was EOS (:418)
    return audio;
}

std::vector<float> TTSEngine::synthesize(const std::string& text, Language lang, const SamplingParams& params) {
    if (!ready_) return {};
    bool ok = false;
    const std::vector<int64_t> ids = wrap_text(text, ok);
    if (!ok) return {};
    return synthesize_tokens(ids, lang, params);
}

std::vector<float> TTSEngine::synthesize_stream(const std::string& text, Language lang, const SamplingParams& params,
                                                const ChunkCallback& on_chunk) {
    if (!ready_) return {};
    bool ok = false;
    const std::vector<int64_t> ids = wrap_text(text, ok);
    if (!ok) return {};
    return run_tokens(ids, lang, params, nullptr, &on_chunk);
}

std::vector<float> TTSEngine::synthesize_tokens(const std::vector<int64_t>& token_ids, Language lang,
                                                const SamplingParams& params) {
    if (!ready_) return {};
    return run_tokens(token_ids, lang, params, nullptr);
}

std::vector<float> TTSEngine::synthesize_clone(const std::string& text, const std::string& ref_audio_path,
                                               Language lang, const SamplingParams& params) {
    if (!ready_) return {};
    if (!has_speaker_encoder_) {
        std::cerr << "[TTSEngine] Speaker encoder not available" << std::endl;
        return {};
    }
    const std::vector<float> spk = extract_speaker_embedding(ref_audio_path);
    if (spk.empty()) {
        std::cerr << "[TTSEngine] Failed to extract speaker embedding" << std::endl;
        return {};
    }
    bool ok = false;
    const std::vector<int64_t> ids = wrap_text(text, ok);
    if (!ok) return {};
    return run_tokens(ids, lang, params, spk.data());
}

std::vector<float> TTSEngine::synthesize_speaker(const std::string& text, Speaker, Language lang,
                                                 const SamplingParams& params) {
    // same stub as the reference (src/tts_onnx.cpp:320-329)
    std::cerr << "[TTSEngine] Preset speakers require CustomVoice model (not yet supported)" << std::endl;
    return synthesize(text, lang, params);
}

std::vector<float> TTSEngine::extract_speaker_embedding(const std::string& audio_path) {
    if (!handle_ || !has_speaker_encoder_) return {};
    int sr = 0;
    std::vector<float> audio = io::read_wav(audio_path, sr);
    if (audio.empty()) {
        std::cerr << "[TTSEngine] Failed to read audio: " << audio_path << std::endl;
        return {};
    }
    if (sr != config::SAMPLE_RATE) audio = io::resample(audio, sr, config::SAMPLE_RATE);

    lqt_info info{};
    lqt_get_info(handle_, &info);
    std::vector<float> out(static_cast<size_t>(info.hidden));
    if (!std::getenv("LEAXER_HOST_MEL")) {
        // log-mel (src/tts_onnx.cpp:347-359, mel config 1024/256/1024/128/0/12000) + the [frames][mels] transposition (:374-380) +
        // the speaker encoder, all on the device: the mel never comes back to the host
        if (lqt_speaker_embed_audio(handle_, audio.data(), static_cast<int64_t>(audio.size()), out.data()) != 0) {
            std::cerr << "[TTSEngine] Synthesis error: " << lqt_last_error(handle_) << std::endl;
            return {};
        }
        return out;
    }
    io::MelConfig mc;                      // host front end ($LEAXER_HOST_MEL, A/B aid): src/tts_onnx.cpp:347-354
    mc.sample_rate = config::SAMPLE_RATE; mc.n_fft = 1024; mc.hop_size = 256; mc.win_size = 1024;
    mc.num_mels = 128; mc.fmin = 0.0f; mc.fmax = 12000.0f;
    io::MelExtractor mel(mc);
    const std::vector<float> m = mel.extract(audio);
    if (m.empty()) {
        std::cerr << "[TTSEngine] Failed to extract mel spectrogram" << std::endl;
        return {};
    }
    // [num_mels][frames] -> [frames][num_mels]  (:374-380)
    const size_t frames = m.size() / 128;
    std::vector<float> mt(m.size());
    for (size_t f = 0; f < frames; ++f)
        for (size_t b = 0; b < 128; ++b) mt[f * 128 + b] = m[b * frames + f];
    if (lqt_speaker_encoder(handle_, mt.data(), static_cast<int32_t>(frames), out.data()) != 0) {
        std::cerr << "[TTSEngine] Synthesis error: " << lqt_last_error(handle_) << std::endl;
        return {};
    }
    return out;
}

} // namespace leaxer_qwen
