// Fuzz driver for the .lqw header parser (csrc/lqw_loader.h: check_lqw / parse_lqw_header), built by tests/test_model_files.py with
// -fsanitize=address,undefined: a corrupt or hostile weight file must be rejected with a reason -- never read outside the header buffer,
// never overflow an offset, never crash (lqw_loader.h:77). No CUDA call is reached: check_lqw parses and validates only.
//   lqw_fuzz GOOD.lqw SCRATCH.lqw ITERATIONS SEED   -> prints "ok <accepted> <rejected>"
#include "lqw_loader.h"

#include <cstdlib>
#include <random>

int main(int argc, char** argv) {
    if (argc < 5) return 2;
    std::vector<unsigned char> good;
    {
        FILE* f = std::fopen(argv[1], "rb");
        if (!f) return 2;
        unsigned char buf[4096];
        size_t n;
        while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) good.insert(good.end(), buf, buf + n);
        std::fclose(f);
    }
    if (!lqt::check_lqw(argv[1]).empty()) return 3;                     // the seed file itself must be accepted
    const int iters = std::atoi(argv[3]);
    std::mt19937 rng(static_cast<unsigned>(std::atoi(argv[4])));
    uint64_t data_start = 0;
    std::memcpy(&data_start, good.data() + 16, 8);
    const size_t hdr = static_cast<size_t>(std::min<uint64_t>(data_start, good.size()));
    long accepted = 0, rejected = 0;
    for (int it = 0; it < iters; ++it) {
        std::vector<unsigned char> m = good;
        const int kind = static_cast<int>(rng() % 6);
        if (kind == 0) {                                                  // a few random bytes inside the header block
            const int n = 1 + static_cast<int>(rng() % 4);
            for (int i = 0; i < n; ++i) m[rng() % hdr] = static_cast<unsigned char>(rng());
        } else if (kind == 1) {                                           // a 32-bit field set to an extreme value
            static const uint32_t ext[] = {0u, 1u, 0x7fffffffu, 0x80000000u, 0xfffffff0u, 0xffffffffu, 0x10000u, 0xffffu};
            const uint32_t v = ext[rng() % 8];
            const size_t p = rng() % (hdr - 4);
            std::memcpy(m.data() + p, &v, 4);
        } else if (kind == 2) {                                           // a 64-bit field set to an extreme value
            static const uint64_t ext[] = {0ull, 8ull, 1ull << 31, 1ull << 40, 0x7fffffffffffffffull, 0xffffffffffffffffull, 0xfffffffffffffff0ull};
            const uint64_t v = ext[rng() % 7];
            const size_t p = rng() % (hdr - 8);
            std::memcpy(m.data() + p, &v, 8);
        } else if (kind == 3) {                                           // truncated anywhere
            m.resize(rng() % (m.size() + 1));
        } else if (kind == 4) {                                           // a 16-bit length field blown up
            const uint16_t v = static_cast<uint16_t>(rng() % 2 ? 0xffffu : rng());
            const size_t p = 24 + rng() % (hdr - 26);
            std::memcpy(m.data() + p, &v, 2);
        } else {                                                          // header bytes shuffled around
            const size_t a = rng() % hdr, b = rng() % hdr, n = std::min<size_t>(1 + rng() % 16, hdr - std::max(a, b));
            for (size_t i = 0; i < n; ++i) std::swap(m[a + i], m[b + i]);
        }
        FILE* f = std::fopen(argv[2], "wb");
        if (!f) return 2;
        if (!m.empty()) std::fwrite(m.data(), 1, m.size(), f);
        std::fclose(f);
        if (lqt::check_lqw(argv[2]).empty()) ++accepted; else ++rejected;
    }
    std::printf("ok %ld %ld\n", accepted, rejected);
    return 0;
}
