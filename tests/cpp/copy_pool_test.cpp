// Hammers jpeg_encoder_b200/csrc/copy_pool.h: many copies of random sizes and thread counts through one pool, every result
// compared with the source; pools are created and destroyed repeatedly, also without ever being used.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../jpeg_encoder_b200/csrc/copy_pool.h"

int main(int argc, char **argv) {
    const int rounds = argc > 1 ? std::atoi(argv[1]) : 2000;
    unsigned seed = 12345;
    auto rnd = [&] { return seed = seed * 1664525u + 1013904223u; };
    std::vector<uint8_t> src(9u << 20), dst(9u << 20);
    for (size_t i = 0; i < src.size(); ++i) src[i] = (uint8_t)(i * 2654435761u >> 13);
    for (int p = 0; p < 8; ++p) {
        jpgb::CopyPool pool;
        if (p == 7) continue; // destroyed unused
        for (int r = 0; r < rounds / 8; ++r) {
            size_t n;
            switch (rnd() % 5) {
            case 0: n = rnd() % 200; break;                   // tiny: parts of zero length
            case 1: n = (rnd() % 64) * 64; break;             // exact multiples of the alignment
            default: n = rnd() % src.size(); break;
            }
            const size_t off = rnd() % (src.size() - n + 1);
            const unsigned threads = rnd() % 6; // 0, 1: caller alone; 5: clamped to 4
            std::fill(dst.begin() + off, dst.begin() + off + n, 0);
            pool.copy(dst.data() + off, src.data() + off, n, threads);
            if (std::memcmp(dst.data() + off, src.data() + off, n) != 0) {
                std::printf("mismatch: n=%zu threads=%u\n", n, threads);
                return 1;
            }
        }
    }
    std::printf("ok\n");
    return 0;
}
