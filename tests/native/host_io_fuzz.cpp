// Fuzz driver for the host parsers that read caller-supplied files (host/io/wav_reader.cpp read_wav: the --ref clip of the clone path,
// src/tts_onnx.cpp:331-340; host/io/tokenizer.cpp load_vocab / load_merges), built by tests/test_host_io.py with
// -fsanitize=address,undefined: mutated and truncated files must be rejected (or read) without an out-of-bounds access or a crash.
//   host_io_fuzz wav GOOD.wav SCRATCH ITERS SEED | host_io_fuzz vocab GOOD.json SCRATCH ITERS SEED | host_io_fuzz merges GOOD.txt SCRATCH ITERS SEED
#include "tokenizer.h"
#include "wav_reader.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

using namespace leaxer_qwen::io;

int main(int argc, char** argv) {
    if (argc < 6) return 2;
    const std::string what = argv[1];
    std::vector<unsigned char> good;
    {
        FILE* f = std::fopen(argv[2], "rb");
        if (!f) return 2;
        unsigned char buf[4096];
        size_t n;
        while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) good.insert(good.end(), buf, buf + n);
        std::fclose(f);
    }
    if (good.size() < 64) return 2;
    const int iters = std::atoi(argv[4]);
    std::mt19937 rng(static_cast<unsigned>(std::atoi(argv[5])));
    const size_t hdr = what == "wav" ? 64 : good.size();              // WAV: the chunk headers are what matters
    long ok = 0, bad = 0;
    for (int it = 0; it < iters; ++it) {
        std::vector<unsigned char> m = good;
        const int kind = static_cast<int>(rng() % 5);
        if (kind == 0) {
            const int n = 1 + static_cast<int>(rng() % 4);
            for (int i = 0; i < n; ++i) m[rng() % hdr] = static_cast<unsigned char>(rng());
        } else if (kind == 1) {
            static const uint32_t ext[] = {0u, 1u, 2u, 0x7fffffffu, 0x80000000u, 0xfffffff0u, 0xffffffffu, 0x10000u};
            const uint32_t v = ext[rng() % 8];
            std::memcpy(m.data() + rng() % (hdr - 4), &v, 4);
        } else if (kind == 2) {
            static const uint16_t ext[] = {0, 1, 2, 3, 8, 24, 32, 0x7fff, 0xffff};
            const uint16_t v = ext[rng() % 9];
            std::memcpy(m.data() + rng() % (hdr - 2), &v, 2);
        } else if (kind == 3) {
            m.resize(rng() % (m.size() + 1));
        } else {
            const size_t a = rng() % m.size(), n = std::min<size_t>(1 + rng() % 64, m.size() - a);
            m.erase(m.begin() + a, m.begin() + a + n);                 // a run of bytes removed
        }
        if (what == "wav") {
            // A data chunk may claim more bytes than the file holds: the reader keeps the reference's behaviour (src/io/wav_reader.cpp: the
            // missing samples stay zero, tests/io_cases.py "truncated_data"), i.e. it allocates what the chunk claims. Bound the claim to
            // 1 MB here so that the campaign tests the parser, not the allocator.
            for (size_t i = 0; i + 8 <= m.size(); ++i)
                if (std::memcmp(m.data() + i, "data", 4) == 0) {
                    uint32_t v; std::memcpy(&v, m.data() + i + 4, 4);
                    if (v > (1u << 20)) { v = (1u << 20) - (v & 0xffu); std::memcpy(m.data() + i + 4, &v, 4); }
                }
        }
        FILE* f = std::fopen(argv[3], "wb");
        if (!f) return 2;
        if (!m.empty()) std::fwrite(m.data(), 1, m.size(), f);
        std::fclose(f);
        bool good_read;
        if (what == "wav") { int sr = 0; good_read = !read_wav(argv[3], sr).empty(); }
        else if (what == "vocab") { good_read = load_vocab(argv[3]); (void)tokenize("hello world 123"); }
        else { good_read = load_merges(argv[3]); (void)tokenize("hello world 123"); }
        if (good_read) ++ok; else ++bad;
    }
    std::printf("ok %ld %ld\n", ok, bad);
    return 0;
}
