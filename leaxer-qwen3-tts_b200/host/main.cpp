// Command-line front end with the reference's flags, defaults, messages and exit codes
// (leaxer-ai/leaxer-qwen3-tts src/main_onnx.cpp:60-192; behaviour: SURVEY.md Appendix D "CLI"):
//   -m/--model DIR  -p/--prompt TEXT  -o/--output PATH  --lang  --ref  --temp  --top-k  --top-p  --max-tokens  -h/--help
// Unknown flags and flags without a value are ignored, -m and -p are required (exit 1 + usage), the model directory
// must exist, the output's parent directory is created, synthesis failure / unwritable output -> exit 1.
// Extensions: --seed N (Philox seed; the reference's sampler is not reproducible), also $LEAXER_SEED;
// --dump-codes PATH (the generated codes, int64 [frames][16] raw, for parity tests); --tokenizer hf|reference (default reference:
// the reference's ASCII-only pre-tokeniser; hf = the published Qwen2 pipeline, io/tokenizer.h; also $LEAXER_TOKENIZER=hf).
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <string>
#include <vector>

#include "io/tokenizer.h"
#include "io/wav_reader.h"
#include "tts_onnx.h"

namespace fs = std::filesystem;
using leaxer_qwen::Language;

namespace {

struct Options {
    const char* model = nullptr;
    const char* prompt = nullptr;
    const char* output = "output.wav";
    const char* lang = "auto";
    const char* ref = nullptr;
    const char* dump_codes = nullptr;
    float temperature = 0.8f, top_p = 0.95f;
    int top_k = 50, max_tokens = 2048;
    const char* tokenizer = nullptr;
    bool have_seed = false;
    unsigned seed = 0;
    bool help = false;
};

void usage(const char* prog) {
    std::printf("Usage: %s [options]\n\n", prog);
    std::printf("Qwen3-TTS ONNX inference\n\n");
    std::printf("Options:\n");
    std::printf("  -m, --model DIR       ONNX model directory (required)\n");
    std::printf("  -p, --prompt TEXT     Text to synthesize (required)\n");
    std::printf("  -o, --output PATH     Output WAV file (default: output.wav)\n");
    std::printf("  --lang LANG           Language: auto, en, zh, ja, ko (default: auto)\n");
    std::printf("  --ref PATH            Reference audio for voice clone (3s WAV)\n");
    std::printf("  --temp FLOAT          Temperature (default: 0.8)\n");
    std::printf("  --top-k N             Top-k sampling (default: 50)\n");
    std::printf("  --top-p FLOAT         Top-p sampling (default: 0.95)\n");
    std::printf("  --max-tokens N        Max tokens (default: 2048)\n");
    std::printf("  -h, --help            Show this help\n");
    std::printf("\nExamples:\n");
    std::printf("  %s -m onnx/onnx_kv_06b -p \"Hello world\" -o hello.wav\n", prog);
    std::printf("  %s -m onnx/onnx_kv_06b -p \"Hello\" --ref voice.wav -o cloned.wav\n", prog);
}

Options parse(int argc, char** argv) {
    Options o;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        const bool has_value = i + 1 < argc;
        if (a == "-h" || a == "--help") { o.help = true; return o; }
        if (!has_value) continue;                                  // a flag without its value is ignored
        if (a == "-m" || a == "--model") o.model = argv[++i];
        else if (a == "-p" || a == "--prompt") o.prompt = argv[++i];
        else if (a == "-o" || a == "--output") o.output = argv[++i];
        else if (a == "--lang") o.lang = argv[++i];
        else if (a == "--ref") o.ref = argv[++i];
        else if (a == "--temp") o.temperature = static_cast<float>(std::atof(argv[++i]));
        else if (a == "--top-k") o.top_k = std::atoi(argv[++i]);
        else if (a == "--top-p") o.top_p = static_cast<float>(std::atof(argv[++i]));
        else if (a == "--max-tokens") o.max_tokens = std::atoi(argv[++i]);
        else if (a == "--dump-codes") o.dump_codes = argv[++i];
        else if (a == "--tokenizer") o.tokenizer = argv[++i];
        else if (a == "--seed") { o.seed = static_cast<unsigned>(std::strtoul(argv[++i], nullptr, 10)); o.have_seed = true; }
    }
    return o;
}

Language language_of(const std::string& s) {
    if (s == "en" || s == "english") return Language::English;
    if (s == "zh" || s == "chinese") return Language::Chinese;
    if (s == "ja" || s == "japanese") return Language::Japanese;
    if (s == "ko" || s == "korean") return Language::Korean;
    return Language::Auto;
}

}  // namespace

int main(int argc, char** argv) {
    const Options o = parse(argc, argv);
    if (o.help) { usage(argv[0]); return 0; }
    if (!o.model || !o.prompt) {
        std::fprintf(stderr, "Error: --model and --prompt are required\n");
        usage(argv[0]);
        return 1;
    }
    if (!fs::exists(o.model)) {
        std::fprintf(stderr, "Error: model directory not found: %s\n", o.model);
        return 1;
    }
    std::printf("Model: %s\n", o.model);
    std::printf("Text: %s\n", o.prompt);
    if (o.ref) std::printf("Reference: %s\n", o.ref);
    std::printf("Language: %s\n", o.lang);
    std::printf("Output: %s\n\n", o.output);

    const fs::path out(o.output);
    if (out.has_parent_path()) fs::create_directories(out.parent_path());

    if (o.tokenizer) leaxer_qwen::io::set_tokenizer_mode(std::string(o.tokenizer) == "hf" ? leaxer_qwen::io::TokenizerMode::HF
                                                                                          : leaxer_qwen::io::TokenizerMode::Reference);
    leaxer_qwen::TTSEngine engine(o.model);
    if (!engine.is_ready()) {
        std::fprintf(stderr, "Error: %s\n", engine.get_error().c_str());
        return 1;
    }
    if (o.have_seed) engine.set_seed(o.seed);

    leaxer_qwen::SamplingParams params;
    params.temperature = o.temperature;
    params.top_k = o.top_k;
    params.top_p = o.top_p;
    params.max_new_tokens = o.max_tokens;

    std::printf("Synthesizing...\n");
    std::vector<float> audio;
    if (o.ref) {
        if (!engine.has_speaker_encoder()) {
            std::fprintf(stderr, "Error: speaker encoder not available for voice clone\n");
            return 1;
        }
        audio = engine.synthesize_clone(o.prompt, o.ref, language_of(o.lang), params);
    } else {
        audio = engine.synthesize(o.prompt, language_of(o.lang), params);
    }
    if (audio.empty()) {
        std::fprintf(stderr, "Error: synthesis failed\n");
        return 1;
    }
    if (o.dump_codes) {
        if (FILE* f = std::fopen(o.dump_codes, "wb")) {
            const std::vector<int64_t>& c = engine.last_codes();
            if (!c.empty()) std::fwrite(c.data(), sizeof(int64_t), c.size(), f);
            std::fclose(f);
        }
    }
    std::printf("Generated %.2f seconds of audio\n", static_cast<float>(audio.size()) / leaxer_qwen::config::SAMPLE_RATE);
    if (leaxer_qwen::io::write_wav_cli(o.output, audio, leaxer_qwen::config::SAMPLE_RATE) != 0) {
        std::fprintf(stderr, "Error: failed to write WAV\n");
        return 1;
    }
    std::printf("Saved to: %s\n", o.output);
    return 0;
}
