// Byte-level BPE tokenizer with the reference's API and observable behaviour
// (leaxer-ai/leaxer-qwen3-tts src/io/tokenizer.h:13-28; behaviour: SURVEY.md Appendix D).
#ifndef LEAXER_QWEN_IO_TOKENIZER_H
#define LEAXER_QWEN_IO_TOKENIZER_H

#include <cstdint>
#include <string>
#include <vector>

namespace leaxer_qwen {
namespace io {

bool load_vocab(const std::string& vocab_path);      // flat JSON object {"token": id, ...}
bool load_merges(const std::string& merges_path);    // one "left right" pair per line, rank = line order
bool is_tokenizer_ready();                           // vocab AND merges loaded
std::vector<int32_t> tokenize(const std::string& text);
std::string token_to_string(int32_t id);             // "" when unknown
int32_t string_to_token(const std::string& token);   // -1 when unknown

// Extension, off by default: the reference's pre-tokeniser is ASCII-only and maps bytes 161-172 / 174-255 to raw
// single bytes, so non-ASCII text (zh/ja/ko) never matches the Qwen vocabulary and degrades to raw byte ids. HF mode is the
// published Qwen2 tokenizer pipeline instead -- the Unicode pre-tokeniser pattern, the GPT-2 byte -> code-point alphabet,
// surrogate pairs in vocab.json -- checked against the HuggingFace `tokenizers` library (tests/test_hf_tokenizer.py). Input is
// expected in NFC (the published pipeline normalises; no normaliser is built in). Specials are still added by id by the engine.
#define LEAXER_HAS_HF_TOKENIZER 1
enum class TokenizerMode { Reference, HF };
void set_tokenizer_mode(TokenizerMode mode);         // process-wide, like the tokenizer itself; also $LEAXER_TOKENIZER=hf
TokenizerMode tokenizer_mode();

} // namespace io
} // namespace leaxer_qwen

#endif
