// B200-native drop-in for the reference engine header (leaxer-ai/leaxer-qwen3-tts src/tts_onnx.h).
//
// The PUBLIC section below is source compatible with the reference: same namespace, constants
// (tts_onnx.h:29-70), enums (:73-93), SamplingParams (:99-105), TTSEngine public methods (:118-164) and
// free helpers (:96, :230-247). What differs is behind it: the seven (+1) Ort::Session members and
// the host-side loops are replaced by one opaque handle of the C-ABI library (include/lqt_b200.h),
// whose sm_100a kernels run prompt assembly, prefill, the frame loop, the sampler and the vocoder
// on the GPU. There is no CPU fallback: without a Blackwell GPU the constructor leaves
// is_ready() == false with an error message, exactly like a reference build without its models.
#ifndef LEAXER_QWEN_TTS_ONNX_H
#define LEAXER_QWEN_TTS_ONNX_H

#include <array>
#include <cstdint>
#include <memory>
#include <optional>
#include <functional>
#include <string>
#include <vector>

struct lqt_engine;   // include/lqt_b200.h

namespace leaxer_qwen {

namespace config {
    // Architecture of the 0.6B graphs the reference was written for. The engine itself reads the
    // real dimensions from the model files (a 1.7B talker loads too); these stay for source compatibility.
    constexpr int HIDDEN_SIZE = 1024;
    constexpr int NUM_LAYERS = 28;
    constexpr int NUM_KV_HEADS = 8;
    constexpr int HEAD_DIM = 128;
    constexpr int VOCAB_SIZE = 3072;
    constexpr int NUM_CODE_GROUPS = 16;
    constexpr int SUBCODE_VOCAB_SIZE = 2048;

    constexpr int64_t TTS_BOS = 151672;
    constexpr int64_t TTS_EOS = 151673;
    constexpr int64_t TTS_PAD = 151671;

    constexpr int64_t IM_START = 151644;
    constexpr int64_t IM_END = 151645;
    constexpr int64_t ASSISTANT = 77091;

    constexpr int64_t CODEC_BOS = 2149;
    constexpr int64_t CODEC_EOS = 2150;
    constexpr int64_t CODEC_PAD = 2148;
    constexpr int64_t CODEC_THINK = 2154;
    constexpr int64_t CODEC_NOTHINK = 2155;
    constexpr int64_t CODEC_THINK_BOS = 2156;
    constexpr int64_t CODEC_THINK_EOS = 2157;

    constexpr int64_t LANG_ENGLISH = 2050;
    constexpr int64_t LANG_CHINESE = 2051;
    constexpr int64_t LANG_JAPANESE = 2052;
    constexpr int64_t LANG_KOREAN = 2053;

    constexpr int MAX_NEW_TOKENS = 2048;
    constexpr float DEFAULT_TEMPERATURE = 0.8f;
    constexpr float DEFAULT_TOP_P = 0.95f;
    constexpr int DEFAULT_TOP_K = 50;
    constexpr int SAMPLE_RATE = 24000;
}

enum class Language { Auto, English, Chinese, Japanese, Korean };

enum class Speaker { None, Serena, Vivian, Uncle_Fu, Dylan, Eric, Ryan, Aiden, Ono_Anna, Sohee };

Speaker parse_speaker(const std::string& name);

struct SamplingParams {
    float temperature = config::DEFAULT_TEMPERATURE;
    float top_p = config::DEFAULT_TOP_P;
    int top_k = config::DEFAULT_TOP_K;
    float repetition_penalty = 1.0f;      // declared by the reference, never read there either
    int max_new_tokens = config::MAX_NEW_TOKENS;
};

class TTSEngine {
public:
    explicit TTSEngine(const std::string& model_dir);
    ~TTSEngine();

    TTSEngine(const TTSEngine&) = delete;
    TTSEngine& operator=(const TTSEngine&) = delete;

    std::vector<float> synthesize(const std::string& text, Language lang = Language::Auto,
                                  const SamplingParams& params = SamplingParams());
    std::vector<float> synthesize_clone(const std::string& text, const std::string& ref_audio_path,
                                        Language lang = Language::Auto,
                                        const SamplingParams& params = SamplingParams());
    std::vector<float> synthesize_speaker(const std::string& text, Speaker speaker,
                                          Language lang = Language::Auto,
                                          const SamplingParams& params = SamplingParams());
    std::vector<float> synthesize_tokens(const std::vector<int64_t>& token_ids, Language lang = Language::Auto,
                                         const SamplingParams& params = SamplingParams());
    std::vector<float> extract_speaker_embedding(const std::string& audio_path);

    bool has_speaker_encoder() const { return has_speaker_encoder_; }
    bool is_ready() const { return ready_; }
    const std::string& get_error() const { return error_msg_; }

    // ---- extensions (not in the reference) ------------------------------------------------------
    // The reference draws from a std::random_device-seeded mt19937 (not reproducible). This engine
    // uses a counter-based Philox stream keyed by (seed, utterance number); the seed defaults to
    // $LEAXER_SEED or, if unset, std::random_device.
    void set_seed(uint32_t seed) { seed_ = seed; }
    uint32_t seed() const { return seed_; }
    // frames (each 16 codes, row-major [frame][codebook]) of the last synthesize* call
    const std::vector<int64_t>& last_codes() const { return last_codes_; }
    // Streaming synthesis: like synthesize(), and on_chunk(pcm, first_sample, n_samples) is called for every 2 s of audio as soon
    // as it is vocoded, while the rest of the utterance is still being generated (the reference delivers everything at the end)
    using ChunkCallback = std::function<void(const float* pcm, int64_t first_sample, int64_t n_samples)>;
    std::vector<float> synthesize_stream(const std::string& text, Language lang, const SamplingParams& params,
                                         const ChunkCallback& on_chunk);

private:
    lqt_engine* handle_ = nullptr;
    bool has_speaker_encoder_ = false;
    bool ready_ = false;
    std::string error_msg_;
    std::string model_dir_;
    uint32_t seed_ = 0;
    uint32_t utterance_ = 0;
    std::vector<int64_t> last_codes_;

    std::vector<int64_t> wrap_text(const std::string& text, bool& ok);
    std::vector<float> run_tokens(const std::vector<int64_t>& ids, Language lang, const SamplingParams& params,
                                  const float* speaker_embed, const ChunkCallback* on_chunk = nullptr);
};

inline int64_t language_to_codec_id(Language lang) {
    switch (lang) {
        case Language::English:  return config::LANG_ENGLISH;
        case Language::Chinese:  return config::LANG_CHINESE;
        case Language::Japanese: return config::LANG_JAPANESE;
        case Language::Korean:   return config::LANG_KOREAN;
        default:                 return 0;
    }
}

// kept for source compatibility: this build has exactly one backend
inline bool is_coreml_enabled() { return false; }

} // namespace leaxer_qwen

#endif // LEAXER_QWEN_TTS_ONNX_H
