// Compiles the C++ mirror (include/jpeg_encoder.hpp) against the C ABI and, on a machine without a
// B200, checks that encode fails loudly (no CPU fallback). With a GPU it encodes a tiny image.
#include <cstdio>
#include <vector>

#include "jpeg_encoder.hpp"

int main() {
    std::vector<uint8_t> out, px(16 * 16 * 3, 128);
    jpeg_encoder::Encoder<jpeg_encoder::VecSink> enc(jpeg_encoder::VecSink{&out}, 90);
    enc.set_sampling_factor(jpeg_encoder::SamplingFactor::F_2_2);
    enc.set_restart_interval(0);
    enc.add_app_segment(15, {'H', 'O', 0});
    try {
        enc.encode(px.data(), px.size(), 16, 16, jpeg_encoder::ColorType::Rgb);
    } catch (const jpeg_encoder::EncodingError &e) {
        std::printf("EncodingError %d: %s\n", e.code, e.what());
        return e.code == JPGB_ERR_CUDA ? 0 : 1;
    }
    // flat grey 16x16 4:2:0: payload 28 A2 8A 00 then EOI
    const size_t n = out.size();
    const bool ok = n > 6 && out[n - 6] == 0x28 && out[n - 5] == 0xA2 && out[n - 4] == 0x8A && out[n - 3] == 0x00 && out[n - 2] == 0xFF && out[n - 1] == 0xD9;
    std::printf("encoded %zu bytes, tail %s\n", n, ok ? "ok" : "BAD");
    return ok ? 0 : 2;
}
