// Re-implementation of the reference tokenizer's behaviour (src/io/tokenizer.cpp), written from its
// behavioural description and checked against the compiled reference on
// random inputs (tests/test_host_io.py). Differences in construction, not in results:
//   * the pre-tokeniser is a hand-written scanner equivalent to the reference's ECMAScript pattern
//     ('s|'t|'re|'ve|'m|'ll|'d| ?[A-Za-z]+|[0-9]+| ?[^\s\w]+|\s+) instead of a std::regex compiled
//     on every call;
//   * the byte -> symbol table is built once; merge ranks live in one string-keyed hash map.
// Quirks that are part of the behaviour and therefore kept: bytes 161-172 / 174-255 map to
// themselves as single raw bytes (not to UTF-8 code points), so non-ASCII text falls back to raw
// byte ids; characters matched by no alternative (e.g. '_') are dropped; the "#version" header line
// of merges.txt is stored as a rank-0 pair.
#include "tokenizer.h"

#include <vector>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

namespace leaxer_qwen {
namespace io {
namespace {

struct Symbols {                          // GPT-2 byte encoder with the reference's raw-byte quirk
    std::string of[256];
    Symbols() {
        int shifted = 0;
        for (int b = 0; b < 256; ++b) {
            const bool direct = (b >= 33 && b <= 126) || (b >= 161 && b <= 172) || (b >= 174 && b <= 255);
            if (direct) {
                of[b] = std::string(1, static_cast<char>(b));
            } else {
                const int cp = 0x100 + shifted++;           // U+0100 + number of non-direct bytes below b
                of[b].push_back(static_cast<char>(0xC0 | (cp >> 6)));
                of[b].push_back(static_cast<char>(0x80 | (cp & 0x3F)));
            }
        }
    }
};
const Symbols& symbols() { static const Symbols s; return s; }

// ---- HF mode (tokenizer.h): the published Qwen2 pipeline ---------------------------------------------------------------
#include "unicode_tables.inc"

void put_utf8(std::string& out, uint32_t cp) {
    if (cp < 0x80) out.push_back(static_cast<char>(cp));
    else if (cp < 0x800) { out.push_back(static_cast<char>(0xC0 | (cp >> 6))); out.push_back(static_cast<char>(0x80 | (cp & 0x3F))); }
    else if (cp < 0x10000) { out.push_back(static_cast<char>(0xE0 | (cp >> 12))); out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
                             out.push_back(static_cast<char>(0x80 | (cp & 0x3F))); }
    else { out.push_back(static_cast<char>(0xF0 | (cp >> 18))); out.push_back(static_cast<char>(0x80 | ((cp >> 12) & 0x3F)));
           out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F))); out.push_back(static_cast<char>(0x80 | (cp & 0x3F))); }
}

struct HfSymbols {                        // GPT-2 bytes_to_unicode: printable bytes keep their code point, the rest go to U+0100..
    std::string of[256];
    HfSymbols() {
        int shifted = 0;
        for (int b = 0; b < 256; ++b) {
            const bool direct = (b >= 33 && b <= 126) || (b >= 161 && b <= 172) || (b >= 174 && b <= 255);
            put_utf8(of[b], direct ? static_cast<uint32_t>(b) : static_cast<uint32_t>(0x100 + shifted++));
        }
    }
};
const HfSymbols& hf_symbols() { static const HfSymbols s; return s; }

template <size_t N>
bool in_ranges(const uint32_t (&tab)[N][2], uint32_t cp) {
    size_t lo = 0, hi = N;
    while (lo < hi) {
        const size_t mid = (lo + hi) / 2;
        if (cp < tab[mid][0]) hi = mid;
        else if (cp > tab[mid][1]) lo = mid + 1;
        else return true;
    }
    return false;
}
inline bool u_letter(uint32_t c) { return in_ranges(kUnicodeLetters, c); }
inline bool u_number(uint32_t c) { return in_ranges(kUnicodeNumbers, c); }
inline bool u_space(uint32_t c) {         // Unicode White_Space (what \s means in the pattern)
    return (c >= 0x9 && c <= 0xD) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) ||
           c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}
inline bool u_newline(uint32_t c) { return c == '\r' || c == '\n'; }
inline bool u_other(uint32_t c) { return !u_space(c) && !u_letter(c) && !u_number(c); }      // [^\s\p{L}\p{N}]

struct CodePoint { uint32_t cp; uint32_t off; };       // code point and its byte offset in the text

// UTF-8 -> code points; a byte that does not start a valid sequence becomes one "other" code point (U+FFFD) of length 1
std::vector<CodePoint> decode_utf8(const std::string& s) {
    std::vector<CodePoint> out;
    const size_t n = s.size();
    size_t i = 0;
    while (i < n) {
        const unsigned char c = static_cast<unsigned char>(s[i]);
        uint32_t cp = 0xFFFD; size_t len = 1;
        auto cont = [&](size_t k) { return i + k < n && (static_cast<unsigned char>(s[i + k]) & 0xC0) == 0x80; };
        if (c < 0x80) cp = c;
        else if ((c & 0xE0) == 0xC0 && c >= 0xC2 && cont(1)) { cp = ((c & 0x1Fu) << 6) | (static_cast<unsigned char>(s[i + 1]) & 0x3Fu); len = 2; }
        else if ((c & 0xF0) == 0xE0 && cont(1) && cont(2)) {
            cp = ((c & 0x0Fu) << 12) | ((static_cast<unsigned char>(s[i + 1]) & 0x3Fu) << 6) | (static_cast<unsigned char>(s[i + 2]) & 0x3Fu); len = 3;
            if (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF)) { cp = 0xFFFD; len = 1; }
        } else if ((c & 0xF8) == 0xF0 && cont(1) && cont(2) && cont(3)) {
            cp = ((c & 0x07u) << 18) | ((static_cast<unsigned char>(s[i + 1]) & 0x3Fu) << 12) | ((static_cast<unsigned char>(s[i + 2]) & 0x3Fu) << 6) |
                 (static_cast<unsigned char>(s[i + 3]) & 0x3Fu); len = 4;
            if (cp < 0x10000 || cp > 0x10FFFF) { cp = 0xFFFD; len = 1; }
        }
        out.push_back(CodePoint{cp, static_cast<uint32_t>(i)});
        i += len;
    }
    return out;
}

// Length (in code points) of the match of the Qwen2 pattern starting exactly at t[i]:
//   (?i:'s|'t|'re|'ve|'m|'ll|'d) | [^\r\n\p{L}\p{N}]?\p{L}+ | \p{N} | ?[^\s\p{L}\p{N}]+[\r\n]* | \s*[\r\n]+ | \s+(?!\S) | \s+
// (leftmost alternative that matches, greedy quantifiers with backtracking). Every code point starts some alternative.
size_t hf_match_at(const std::vector<CodePoint>& t, size_t i) {
    const size_t n = t.size();
    const uint32_t c = t[i].cp;
    auto lower = [](uint32_t x) { return (x >= 'A' && x <= 'Z') ? x + 32 : x; };
    if (c == '\'' && i + 1 < n) {
        const uint32_t a = lower(t[i + 1].cp), b = (i + 2 < n) ? lower(t[i + 2].cp) : 0;
        if (a == 's' || a == 't') return 2;
        if (a == 'r' && b == 'e') return 3;
        if (a == 'v' && b == 'e') return 3;
        if (a == 'm') return 2;
        if (a == 'l' && b == 'l') return 3;
        if (a == 'd') return 2;
    }
    {   // [^\r\n\p{L}\p{N}]?\p{L}+
        size_t j = i;
        if (!u_newline(c) && !u_letter(c) && !u_number(c)) j = i + 1;          // the optional prefix (greedy: tried first)
        size_t k = j;
        while (k < n && u_letter(t[k].cp)) ++k;
        if (k > j) return k - i;
        // (without the prefix the first code point would have to be a letter, which the branch above already covers)
    }
    if (u_number(c)) return 1;                                                // \p{N}
    {   //  ?[^\s\p{L}\p{N}]+[\r\n]*
        size_t j = i + ((c == ' ') ? 1 : 0);
        size_t k = j;
        while (k < n && u_other(t[k].cp)) ++k;
        if (k > j) { while (k < n && u_newline(t[k].cp)) ++k; return k - i; }
    }
    if (u_space(c)) {
        size_t e = i;
        while (e < n && u_space(t[e].cp)) ++e;                                // the whitespace run [i, e)
        size_t last_nl = n;
        for (size_t k = i; k < e; ++k) if (u_newline(t[k].cp)) last_nl = k;
        if (last_nl != n) return last_nl + 1 - i;                             // \s*[\r\n]+ : up to the last newline of the run
        if (e == n) return e - i;                                             // \s+(?!\S) at the end of the text
        if (e - i >= 2) return e - i - 1;                                     // \s+(?!\S): leaves one space for the next word
        return e - i;                                                         // \s+
    }
    return 1;
}

// ---- Unicode normalisation form C (UAX #15): the first stage of the published Qwen2 pipeline ("normalizer": NFC) -------------
#include "unicode_nfc_tables.inc"

uint32_t nfc_ccc(uint32_t cp) {
    if (cp < 0x300) return 0;
    size_t lo = 0, hi = sizeof(kNfcCcc) / sizeof(kNfcCcc[0]);
    while (lo < hi) {
        const size_t mid = (lo + hi) / 2;
        if (cp < kNfcCcc[mid][0]) hi = mid;
        else if (cp > kNfcCcc[mid][1]) lo = mid + 1;
        else return kNfcCcc[mid][2];
    }
    return 0;
}
// Hangul syllables are (de)composed arithmetically (UAX #15, section 3.12 of the standard)
constexpr uint32_t kSBase = 0xAC00, kLBase = 0x1100, kVBase = 0x1161, kTBase = 0x11A7, kLCount = 19, kVCount = 21, kTCount = 28,
                   kNCount = kVCount * kTCount, kSCount = kLCount * kNCount;
void nfc_decompose(uint32_t cp, std::vector<uint32_t>& out) {       // full canonical decomposition of one code point
    if (cp < 0xC0) { out.push_back(cp); return; }
    if (cp >= kSBase && cp < kSBase + kSCount) {
        const uint32_t s = cp - kSBase, t = s % kTCount;
        out.push_back(kLBase + s / kNCount);
        out.push_back(kVBase + (s % kNCount) / kTCount);
        if (t) out.push_back(kTBase + t);
        return;
    }
    size_t lo = 0, hi = sizeof(kNfcDecomp) / sizeof(kNfcDecomp[0]);
    while (lo < hi) {
        const size_t mid = (lo + hi) / 2;
        if (kNfcDecomp[mid][0] < cp) lo = mid + 1; else hi = mid;
    }
    if (lo < sizeof(kNfcDecomp) / sizeof(kNfcDecomp[0]) && kNfcDecomp[lo][0] == cp) {
        nfc_decompose(kNfcDecomp[lo][1], out);                      // (the mappings are one step: the first part may decompose further)
        if (kNfcDecomp[lo][2]) nfc_decompose(kNfcDecomp[lo][2], out);
        return;
    }
    out.push_back(cp);
}
uint32_t nfc_compose_pair(uint32_t a, uint32_t b) {                 // primary composite of (a, b), or 0
    if (a >= kLBase && a < kLBase + kLCount && b >= kVBase && b < kVBase + kVCount)
        return kSBase + ((a - kLBase) * kVCount + (b - kVBase)) * kTCount;
    if (a >= kSBase && a < kSBase + kSCount && (a - kSBase) % kTCount == 0 && b > kTBase && b < kTBase + kTCount)
        return a + (b - kTBase);
    size_t lo = 0, hi = sizeof(kNfcComp) / sizeof(kNfcComp[0]);
    while (lo < hi) {
        const size_t mid = (lo + hi) / 2;
        if (kNfcComp[mid][0] < a || (kNfcComp[mid][0] == a && kNfcComp[mid][1] < b)) lo = mid + 1; else hi = mid;
    }
    if (lo < sizeof(kNfcComp) / sizeof(kNfcComp[0]) && kNfcComp[lo][0] == a && kNfcComp[lo][1] == b) return kNfcComp[lo][2];
    return 0;
}
// NFC of a run of code points: canonical decomposition, canonical ordering, canonical composition
std::vector<uint32_t> nfc_run(const std::vector<uint32_t>& in) {
    std::vector<uint32_t> d;
    d.reserve(in.size() + 8);
    for (uint32_t cp : in) nfc_decompose(cp, d);
    std::vector<uint32_t> cc(d.size());
    for (size_t i = 0; i < d.size(); ++i) cc[i] = nfc_ccc(d[i]);
    for (size_t i = 1; i < d.size(); ++i) {                          // stable insertion sort inside every run of non-starters
        if (cc[i] == 0) continue;
        size_t j = i;
        while (j > 0 && cc[j - 1] > cc[j]) { std::swap(cc[j - 1], cc[j]); std::swap(d[j - 1], d[j]); --j; }
    }
    std::vector<uint32_t> out;
    out.reserve(d.size());
    size_t starter = static_cast<size_t>(-1);
    uint32_t last_cc = 0;
    for (size_t i = 0; i < d.size(); ++i) {
        // d[i] combines with the last starter unless something stands between them whose class is 0 or >= its own ("blocked")
        if (starter != static_cast<size_t>(-1) && (out.size() - 1 == starter || last_cc < cc[i])) {
            const uint32_t c = nfc_compose_pair(out[starter], d[i]);
            if (c) { out[starter] = c; continue; }
        }
        if (cc[i] == 0) starter = out.size();
        last_cc = cc[i];
        out.push_back(d[i]);
    }
    return out;
}
// NFC of UTF-8 text. Bytes that are not valid UTF-8 pass through unchanged (and separate the runs); text without any code point
// >= U+0300 (lead bytes < 0xCC) is already normalised and returned as it is.
std::string nfc_utf8(const std::string& text) {
    bool plain = true;
    for (unsigned char c : text) if (c >= 0xCC) { plain = false; break; }
    if (plain) return text;
    const std::vector<CodePoint> t = decode_utf8(text);
    std::string out;
    out.reserve(text.size());
    std::vector<uint32_t> run;
    auto flush = [&]() {
        if (run.empty()) return;
        for (uint32_t cp : nfc_run(run)) put_utf8(out, cp);
        run.clear();
    };
    for (size_t i = 0; i < t.size(); ++i) {
        const size_t len = ((i + 1 < t.size()) ? t[i + 1].off : text.size()) - t[i].off;
        if (t[i].cp == 0xFFFD && len == 1) { flush(); out.push_back(text[t[i].off]); }      // an invalid byte (a real U+FFFD is 3 bytes)
        else run.push_back(t[i].cp);
    }
    flush();
    return out;
}

TokenizerMode g_mode = (std::getenv("LEAXER_TOKENIZER") && std::string(std::getenv("LEAXER_TOKENIZER")) == "hf") ? TokenizerMode::HF
                                                                                                              : TokenizerMode::Reference;

inline bool is_space(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
inline bool is_alpha(unsigned char c) { return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'); }
inline bool is_digit(unsigned char c) { return c >= '0' && c <= '9'; }
inline bool is_word(unsigned char c) { return is_alpha(c) || is_digit(c) || c == '_'; }
inline bool is_other(unsigned char c) { return !is_space(c) && !is_word(c); }     // [^\s\w]

// length of the match of the reference pattern starting exactly at s[i], 0 if none
size_t match_at(const std::string& s, size_t i) {
    const size_t n = s.size();
    const unsigned char c = static_cast<unsigned char>(s[i]);
    if (c == '\'' && i + 1 < n) {                                    // contractions, in pattern order
        const char a = s[i + 1];
        const char b = (i + 2 < n) ? s[i + 2] : '\0';
        if (a == 's' || a == 't') return 2;
        if (a == 'r' && b == 'e') return 3;
        if (a == 'v' && b == 'e') return 3;
        if (a == 'm') return 2;
        if (a == 'l' && b == 'l') return 3;
        if (a == 'd') return 2;
    }
    {   // " ?[A-Za-z]+"
        size_t j = i + ((c == ' ') ? 1 : 0);
        size_t k = j;
        while (k < n && is_alpha(static_cast<unsigned char>(s[k]))) ++k;
        if (k > j) return k - i;
    }
    if (is_digit(c)) {                                               // "[0-9]+"
        size_t k = i;
        while (k < n && is_digit(static_cast<unsigned char>(s[k]))) ++k;
        return k - i;
    }
    {   // " ?[^\s\w]+"   (a leading ' ' that is not followed by such a character backtracks to no space:
        //                then s[i] == ' ' itself is \s and the alternative fails)
        size_t j = i + ((c == ' ') ? 1 : 0);
        size_t k = j;
        while (k < n && is_other(static_cast<unsigned char>(s[k]))) ++k;
        if (k > j) return k - i;
    }
    if (is_space(c)) {                                               // "\s+"
        size_t k = i;
        while (k < n && is_space(static_cast<unsigned char>(s[k]))) ++k;
        return k - i;
    }
    return 0;
}

class Bpe {
public:
    bool read_vocab(const std::string& path);
    bool read_merges(const std::string& path);
    bool vocab_ok() const { return vocab_ok_; }
    bool merges_ok() const { return merges_ok_; }
    std::vector<int32_t> encode(const std::string& text) const;
    std::string text_of(int32_t id) const { auto it = by_id_.find(id); return it == by_id_.end() ? std::string() : it->second; }
    int32_t id_of(const std::string& tok) const { auto it = by_text_.find(tok); return it == by_text_.end() ? -1 : it->second; }

private:
    static std::string pair_key(const std::string& a, const std::string& b) {
        std::string k = std::to_string(a.size());
        k.push_back(':'); k += a; k += b;
        return k;
    }
    void merge_chunk(const std::string& chunk, std::vector<std::string>& out, const std::string (&sym)[256]) const;
    void emit(const std::string& chunk, std::vector<std::string>& parts, const std::string (&sym)[256], std::vector<int32_t>& ids) const;

    bool vocab_ok_ = false, merges_ok_ = false;
    std::unordered_map<std::string, int32_t> by_text_;
    std::unordered_map<int32_t, std::string> by_id_;
    std::unordered_map<std::string, int> rank_;
    size_t n_merges_ = 0;
};

int hex_value(char c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}

bool Bpe::read_vocab(const std::string& path) {
    by_text_.clear(); by_id_.clear();
    vocab_ok_ = false;
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { std::fprintf(stderr, "Failed to open vocab file: %s\n", path.c_str()); return false; }
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (sz <= 0 || sz > 100L * 1024 * 1024) { std::fprintf(stderr, "Invalid file size: %ld\n", sz); std::fclose(f); return false; }
    std::string buf(static_cast<size_t>(sz), '\0');
    const size_t got = std::fread(&buf[0], 1, buf.size(), f);
    std::fclose(f);
    if (got != buf.size()) { std::fprintf(stderr, "Failed to read file\n"); return false; }

    const size_t n = buf.size();
    size_t p = 0;
    auto skip_ws = [&]() { while (p < n && std::isspace(static_cast<unsigned char>(buf[p]))) ++p; };
    skip_ws();
    if (p >= n || buf[p] != '{') { std::fprintf(stderr, "Expected '{' at start of JSON\n"); return false; }
    ++p;
    int count = 0;
    for (;;) {
        skip_ws();
        if (p >= n || buf[p] == '}') break;
        if (buf[p] == ',') { ++p; continue; }
        if (buf[p] != '"') { std::fprintf(stderr, "Expected '\"' at position %zu\n", p); return false; }
        ++p;
        std::string key;
        bool has_surrogate = false;
        while (p < n && buf[p] != '"') {
            char ch = buf[p];
            if (ch != '\\') { key.push_back(ch); ++p; continue; }
            if (++p >= n) { std::fprintf(stderr, "Unexpected end of file in escape sequence\n"); return false; }
            ch = buf[p];
            if (ch == 'n') key.push_back('\n');
            else if (ch == 't') key.push_back('\t');
            else if (ch == 'r') key.push_back('\r');
            else if (ch == 'u') {
                if (p + 4 >= n) { std::fprintf(stderr, "Invalid unicode escape\n"); return false; }
                int cp = 0;
                for (int i = 1; i <= 4; ++i) {
                    const int d = hex_value(buf[p + i]);
                    if (d < 0) { std::fprintf(stderr, "Invalid hex digit in unicode escape\n"); return false; }
                    cp = cp * 16 + d;
                }
                if (cp < 0x80) key.push_back(static_cast<char>(cp));
                else if (cp < 0x800) { key.push_back(static_cast<char>(0xC0 | (cp >> 6))); key.push_back(static_cast<char>(0x80 | (cp & 0x3F))); }
                else { key.push_back(static_cast<char>(0xE0 | (cp >> 12))); key.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
                       key.push_back(static_cast<char>(0x80 | (cp & 0x3F))); }          // BMP only, surrogates not joined
                if (cp >= 0xD800 && cp <= 0xDFFF) has_surrogate = true;
                p += 4;
            } else key.push_back(ch);                                                    // '\\', '"' and anything else: literal
            ++p;
        }
        if (p >= n) { std::fprintf(stderr, "Unexpected end of file in token string\n"); return false; }
        ++p;
        skip_ws();
        if (p >= n || buf[p] != ':') { std::fprintf(stderr, "Expected ':' after token\n"); return false; }
        ++p;
        skip_ws();
        if (p >= n || !std::isdigit(static_cast<unsigned char>(buf[p]))) { std::fprintf(stderr, "Expected digit for token ID\n"); return false; }
        int32_t id = 0;
        while (p < n && std::isdigit(static_cast<unsigned char>(buf[p]))) { id = id * 10 + (buf[p] - '0'); ++p; }
        by_text_[key] = id;
        by_id_[id] = key;
        if (has_surrogate) {
            // HF mode: the same token with its surrogate pairs joined into 4-byte UTF-8 (an extra key; a reference-mode lookup can never
            // produce it, because there every byte >= 161 is a raw single byte)
            std::string joined;
            for (size_t q = 0; q < key.size();) {
                const unsigned char b0 = static_cast<unsigned char>(key[q]);
                if (b0 == 0xED && q + 5 < key.size() && (static_cast<unsigned char>(key[q + 1]) & 0xF0) == 0xA0 &&
                    static_cast<unsigned char>(key[q + 3]) == 0xED && (static_cast<unsigned char>(key[q + 4]) & 0xF0) == 0xB0) {
                    const uint32_t hi = 0xD000u | ((static_cast<unsigned char>(key[q + 1]) & 0x3Fu) << 6) | (static_cast<unsigned char>(key[q + 2]) & 0x3Fu);
                    const uint32_t lo = 0xD000u | ((static_cast<unsigned char>(key[q + 4]) & 0x3Fu) << 6) | (static_cast<unsigned char>(key[q + 5]) & 0x3Fu);
                    put_utf8(joined, 0x10000u + ((hi - 0xD800u) << 10) + (lo - 0xDC00u));
                    q += 6;
                } else { joined.push_back(key[q]); ++q; }
            }
            by_text_[joined] = id;
        }
        if (++count % 10000 == 0) std::fprintf(stderr, "Loaded %d tokens...\n", count);
    }
    std::fprintf(stderr, "Successfully loaded %d tokens\n", count);
    if (by_text_.empty()) return false;
    vocab_ok_ = true;
    return true;
}

bool Bpe::read_merges(const std::string& path) {
    rank_.clear(); n_merges_ = 0;
    FILE* f = std::fopen(path.c_str(), "r");
    if (!f) { std::fprintf(stderr, "Failed to open merges file: %s\n", path.c_str()); return false; }
    char line[1024];
    int rank = 0;
    while (std::fgets(line, sizeof(line), f)) {
        size_t len = std::strlen(line);
        while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = '\0';
        if (len == 0) continue;
        char* sp = std::strchr(line, ' ');
        if (!sp) { std::fprintf(stderr, "Invalid merge line (no space): %s\n", line); continue; }
        *sp = '\0';
        rank_[pair_key(line, sp + 1)] = rank;          // a repeated pair keeps its LAST rank
        ++rank;
        if (++n_merges_ % 10000 == 0) std::fprintf(stderr, "Loaded %zu merge rules...\n", n_merges_);
    }
    std::fclose(f);
    std::fprintf(stderr, "Successfully loaded %zu merge rules\n", n_merges_);
    merges_ok_ = true;                                  // set even when the file was empty, like the reference
    return n_merges_ > 0;
}

void Bpe::merge_chunk(const std::string& chunk, std::vector<std::string>& out, const std::string (&sym)[256]) const {
    out.clear();
    for (unsigned char c : chunk) out.push_back(sym[c]);
    while (out.size() > 1) {
        int best = INT_MAX;
        size_t at = 0;
        for (size_t i = 0; i + 1 < out.size(); ++i) {
            auto it = rank_.find(pair_key(out[i], out[i + 1]));
            if (it != rank_.end() && it->second < best) { best = it->second; at = i; }    // leftmost on equal rank
        }
        if (best == INT_MAX) break;
        out[at] += out[at + 1];
        out.erase(out.begin() + static_cast<long>(at) + 1);
    }
}

void Bpe::emit(const std::string& chunk, std::vector<std::string>& parts, const std::string (&sym)[256], std::vector<int32_t>& ids) const {
    if (merges_ok_) merge_chunk(chunk, parts, sym);
    else { parts.clear(); for (char c : chunk) parts.push_back(std::string(1, c)); }
    for (const std::string& tok : parts) {
        auto it = by_text_.find(tok);
        if (it != by_text_.end()) ids.push_back(it->second);
        else for (unsigned char c : tok) ids.push_back(c);         // unknown symbol: its byte values
    }
}

std::vector<int32_t> Bpe::encode(const std::string& text) const {
    std::vector<int32_t> ids;
    if (text.empty()) return ids;
    if (!vocab_ok_) {                                   // no vocab: raw byte values
        for (unsigned char c : text) ids.push_back(c);
        return ids;
    }
    std::vector<std::string> parts;
    if (g_mode == TokenizerMode::HF) {
        const std::string norm = nfc_utf8(text);         // "normalizer": NFC, then the pre-tokeniser pattern, then byte-level BPE
        const std::vector<CodePoint> t = decode_utf8(norm);
        size_t i = 0;
        while (i < t.size()) {
            const size_t m = hf_match_at(t, i);
            const size_t b0 = t[i].off, b1 = (i + m < t.size()) ? t[i + m].off : norm.size();
            emit(norm.substr(b0, b1 - b0), parts, hf_symbols().of, ids);
            i += m;
        }
        return ids;
    }
    size_t i = 0;
    while (i < text.size()) {
        const size_t m = match_at(text, i);
        if (m == 0) { ++i; continue; }                  // matched by no alternative: dropped
        const std::string chunk = text.substr(i, m);
        i += m;
        emit(chunk, parts, symbols().of, ids);
    }
    return ids;
}

Bpe& instance() { static Bpe t; return t; }

}  // namespace

bool load_vocab(const std::string& vocab_path) { return instance().read_vocab(vocab_path); }
bool load_merges(const std::string& merges_path) { return instance().read_merges(merges_path); }
bool is_tokenizer_ready() { return instance().vocab_ok() && instance().merges_ok(); }
std::vector<int32_t> tokenize(const std::string& text) { return instance().encode(text); }
std::string token_to_string(int32_t id) { return instance().text_of(id); }
int32_t string_to_token(const std::string& token) { return instance().id_of(token); }
std::string normalize_nfc(const std::string& text) { return nfc_utf8(text); }
void set_tokenizer_mode(TokenizerMode mode) { g_mode = mode; }
TokenizerMode tokenizer_mode() { return g_mode; }

} // namespace io
} // namespace leaxer_qwen
