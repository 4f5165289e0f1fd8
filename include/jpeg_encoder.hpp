// jpeg_encoder.hpp -- header-only C++ mirror of the reference crate's public encode API over the
// C ABI in jpegenc_b200.h. The reference is compiled code (Rust); no Rust toolchain exists in the
// build image, so this C++ class is the compiled-language host side. Names, argument meaning and
// error behaviour follow jpeg_encoder::Encoder (/root/reference/src/encoder.rs:213-515); the Rust
// shim a maintainer would add instead is in INTEGRATION.md and rust/.
//
//   jpeg_encoder::Encoder<VecSink> enc(VecSink{&bytes}, 90);
//   enc.set_sampling_factor(jpeg_encoder::SamplingFactor::F_2_2);
//   enc.encode(pixels, len, 1920, 1080, jpeg_encoder::ColorType::Rgb);
#pragma once

#include <cstdint>
#include <exception>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "jpegenc_b200.h"

namespace jpeg_encoder {

enum class ColorType : uint8_t { Luma = 0, Rgb, Rgba, Bgr, Bgra, Ycbcr, Cmyk, CmykAsYcck, Ycck }; // encoder.rs:72-99

enum class SamplingFactor : uint8_t { // encoder.rs:120-153
    F_1_1 = 1 << 4 | 1, F_2_1 = 2 << 4 | 1, F_1_2 = 1 << 4 | 2, F_2_2 = 2 << 4 | 2,
    F_4_1 = 4 << 4 | 1, F_4_2 = 4 << 4 | 2, F_1_4 = 1 << 4 | 4, F_2_4 = 2 << 4 | 4,
    R_4_4_4 = 0x80 | 1 << 4 | 1, R_4_4_0 = 0x80 | 1 << 4 | 2, R_4_4_1 = 0x80 | 1 << 4 | 4, R_4_2_2 = 0x80 | 2 << 4 | 1,
    R_4_2_0 = 0x80 | 2 << 4 | 2, R_4_2_1 = 0x80 | 2 << 4 | 4, R_4_1_1 = 0x80 | 4 << 4 | 1, R_4_1_0 = 0x80 | 4 << 4 | 2,
};

enum class PixelDensityUnit : uint8_t { PixelAspectRatio = 0, Inches = 1, Centimeters = 2 }; // writer.rs:47-59
struct PixelDensity {                                                                          // writer.rs:16-45
    std::pair<uint16_t, uint16_t> density{1, 1};
    PixelDensityUnit unit = PixelDensityUnit::PixelAspectRatio;
    static PixelDensity dpi(uint16_t d) { return PixelDensity{{d, d}, PixelDensityUnit::Inches}; }
};

// quantization.rs:8-40. `Custom` carries 64 values in natural order.
struct QuantizationTableType {
    uint8_t kind = JPGB_QT_DEFAULT;
    uint16_t custom[64] = {};
    static QuantizationTableType Preset(uint8_t k) { QuantizationTableType t; t.kind = k; return t; }
    static QuantizationTableType Custom(const uint16_t (&v)[64]) {
        QuantizationTableType t;
        t.kind = JPGB_QT_CUSTOM;
        for (int i = 0; i < 64; ++i) t.custom[i] = v[i];
        return t;
    }
};

// error.rs:6-28
struct EncodingError : std::runtime_error {
    int code;
    EncodingError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

// One device context per host thread (jpgb_encoder); shared by every Encoder created on that thread.
inline jpgb_encoder *thread_context(int device = 0) {
    thread_local jpgb_encoder *ctx = nullptr;
    if (!ctx && jpgb_encoder_create(device, nullptr, &ctx) != JPGB_OK)
        throw EncodingError(JPGB_ERR_CUDA, "no usable sm_100 GPU (the encode path has no CPU fallback)");
    return ctx;
}

// W needs `void write_all(const uint8_t*, size_t)` (JfifWrite, writer.rs:76-82); it may throw.
template <typename W>
class Encoder {
  public:
    Encoder(W w, uint8_t quality) : w_(std::move(w)) { jpgb_params_default(&p_, quality); } // encoder.rs:239-275

    void set_density(PixelDensity d) { p_.density_unit = (uint8_t)d.unit; p_.density_x = d.density.first; p_.density_y = d.density.second; }
    PixelDensity density() const { return PixelDensity{{p_.density_x, p_.density_y}, (PixelDensityUnit)p_.density_unit}; }
    void set_sampling_factor(SamplingFactor s) { p_.sampling = (uint8_t)s; }
    SamplingFactor sampling_factor() const { return (SamplingFactor)p_.sampling; }
    void set_quantization_tables(const QuantizationTableType &luma, const QuantizationTableType &chroma) {
        const QuantizationTableType *t[2] = {&luma, &chroma};
        for (int i = 0; i < 2; ++i) {
            p_.qtable_kind[i] = t[i]->kind;
            for (int k = 0; k < 64; ++k) p_.qtable_custom[i][k] = t[i]->custom[k];
        }
    }
    void set_progressive(bool on) { p_.progressive_scans = on ? 4 : 0; }              // encoder.rs:317-319
    void set_progressive_scans(uint8_t scans) {                                       // :328-335 (the reference panics)
        if (scans < 2 || scans > 64) throw std::invalid_argument("Invalid number of scans");
        p_.progressive_scans = scans;
    }
    int progressive_scans() const { return p_.progressive_scans ? p_.progressive_scans : -1; }
    void set_restart_interval(uint16_t interval) { p_.restart_interval = interval; }  // :345-347 (0 = off)
    void set_optimized_huffman_tables(bool on) { p_.optimize_huffman = on ? 1 : 0; }  // :357-359
    bool optimized_huffman_tables() const { return p_.optimize_huffman != 0; }

    void add_app_segment(uint8_t nr, std::vector<uint8_t> data) {                     // :374-383
        if (nr == 0 || nr > 15) throw EncodingError(JPGB_ERR_INVALID_APP_SEGMENT, "Invalid app segment number");
        if (data.size() > 65533) throw EncodingError(JPGB_ERR_APP_SEGMENT_TOO_LARGE, "App segment exceeds maximum allowed data length of 65533");
        apps_.emplace_back(nr, std::move(data));
    }
    void add_icc_profile(const uint8_t *data, size_t len) {                            // :392-417
        static const char marker[12] = {'I', 'C', 'C', '_', 'P', 'R', 'O', 'F', 'I', 'L', 'E', 0};
        const size_t max_chunk = 65535 - 2 - 12 - 2, n = (len + max_chunk - 1) / max_chunk;
        if (n >= 255) throw EncodingError(JPGB_ERR_BAD_PARAMS, "ICC profile exceeds maximum allowed data length");
        for (size_t i = 0; i < n; ++i) {
            std::vector<uint8_t> c(marker, marker + 12);
            c.push_back((uint8_t)(i + 1));
            c.push_back((uint8_t)n);
            const size_t a = i * max_chunk, b = a + max_chunk < len ? a + max_chunk : len;
            c.insert(c.end(), data + a, data + b);
            add_app_segment(2, std::move(c));
        }
    }
    void add_exif_metadata(const uint8_t *data, size_t len) {                          // :426-435
        std::vector<uint8_t> f = {0x45, 0x78, 0x69, 0x66, 0x00, 0x00};
        f.insert(f.end(), data, data + len);
        add_app_segment(1, std::move(f));
    }

    // Encoder::encode, encoder.rs:440-503. Consumes the encoder's configuration for one image.
    void encode(const uint8_t *data, size_t len, uint16_t width, uint16_t height, ColorType color) {
        std::vector<jpgb_app_segment> segs;
        for (auto &a : apps_) segs.push_back(jpgb_app_segment{a.first, a.second.data(), (uint32_t)a.second.size()});
        p_.width = width;
        p_.height = height;
        p_.color_type = (uint8_t)color;
        p_.n_app = (uint32_t)segs.size();
        p_.apps = segs.data();
        jpgb_encoder *ctx = thread_context();
        const int rc = jpgb_encode_to_sink(ctx, &p_, data, len, &Encoder::sink, this);
        p_.apps = nullptr;
        p_.n_app = 0;
        if (pending_) std::rethrow_exception(std::exchange(pending_, nullptr));
        if (rc != JPGB_OK) throw EncodingError(rc, jpgb_last_error(ctx));
    }

  private:
    static int sink(void *user, const uint8_t *buf, size_t len) {
        auto *self = static_cast<Encoder *>(user);
        try {
            self->w_.write_all(buf, len);
            return 0;
        } catch (...) {
            self->pending_ = std::current_exception();
            return 1;
        }
    }
    W w_;
    jpgb_params p_{};
    std::vector<std::pair<uint8_t, std::vector<uint8_t>>> apps_;
    std::exception_ptr pending_ = nullptr;
};

struct VecSink { // Vec<u8> sink (writer.rs:91-97)
    std::vector<uint8_t> *out;
    void write_all(const uint8_t *b, size_t n) { out->insert(out->end(), b, b + n); }
};

} // namespace jpeg_encoder
