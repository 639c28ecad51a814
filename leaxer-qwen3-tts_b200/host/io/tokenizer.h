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

} // namespace io
} // namespace leaxer_qwen

#endif
