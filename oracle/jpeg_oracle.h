/*
 * oracle/jpeg_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the encode path of vstroebel/jpeg-encoder v0.7.0.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call this. The product (jpeg_encoder_b200/csrc) never does.
 *
 * Parity status: the reference is Rust and no Rust toolchain exists in this image, so
 * the reference itself cannot be run here. The restatement is pinned against every
 * known-answer vector the reference's own tests hold for this path (fdct.rs:249-274,
 * image_buffer.rs:326-421, quantization.rs:314-338, encoder.rs:1286-1300,
 * lib.rs:417,496,525) and against the hand-derived stream KATs of SURVEY.md section 0.
 * File-level bytes are NOT pinned by the reference's own tests (it holds no golden JPEGs);
 * see DESIGN.md "Oracle".
 */
#ifndef JPEG_ORACLE_H
#define JPEG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ColorType, src/encoder.rs:72-99 */
enum {
    ORC_LUMA = 0, ORC_RGB = 1, ORC_RGBA = 2, ORC_BGR = 3, ORC_BGRA = 4,
    ORC_YCBCR = 5, ORC_CMYK = 6, ORC_CMYK_AS_YCCK = 7, ORC_YCCK = 8
};

typedef struct {
    uint8_t nr;           /* APPn number 1..15 */
    const uint8_t *data;
    uint32_t len;         /* <= 65533 */
} orc_app_segment;

typedef struct {
    uint16_t width, height;
    uint8_t color_type;            /* ORC_* */
    uint8_t quality;               /* Encoder::new(w, quality) */
    uint8_t sampling;              /* (h<<4)|v, alias bit 0x80 ignored; src/encoder.rs:120-176 */
    uint8_t qtable_kind[2];        /* 0..8 = QuantizationTableType preset index, 9 = Custom */
    uint16_t qtable_custom[2][64]; /* natural order, used when kind==9 */
    uint8_t progressive_scans;     /* 0 = off, else 2..=64 */
    uint8_t optimize_huffman;
    uint16_t restart_interval;     /* 0 = off */
    uint8_t density_unit;          /* 0 aspect ratio, 1 inches, 2 cm */
    uint16_t density_x, density_y;
    uint32_t n_app;
    const orc_app_segment *apps;
} orc_params;

/* error codes mirroring src/error.rs */
enum { ORC_OK = 0, ORC_BAD_IMAGE_DATA = 1, ORC_ZERO_DIMENSIONS = 2, ORC_BAD_PARAMS = 3, ORC_NOMEM = 4 };

/* Encoder::encode, src/encoder.rs:440-567. *out is malloc'ed; release with orc_free. */
int orc_encode(const orc_params *p, const uint8_t *pixels, size_t len, uint8_t **out, size_t *out_len);
void orc_free(void *p);

/* Quantized zig-zag blocks of every component over the MCU-padded block grid
 * (comp c: (mcu_rows*V_c) x (mcu_cols*H_c) blocks, raster order, 64 i16 each).
 * Same arithmetic as encode_blocks / the interleaved walk (src/encoder.rs:759-789, 977-1056);
 * used to check the colour+DCT+quant kernel on its own.
 * blocks[c] is malloc'ed (release with orc_free); n_blocks[c] receives the count. */
int orc_coefficients(const orc_params *p, const uint8_t *pixels, size_t len,
                     int16_t *blocks[4], uint32_t n_blocks[4]);

/* unit-level functions for the reference's known-answer tests */
void orc_rgb_to_ycbcr(uint8_t r, uint8_t g, uint8_t b, uint8_t out[3]);     /* image_buffer.rs:9-31 */
void orc_fdct(int16_t block[64]);                                             /* fdct.rs:107-238 */
void orc_fdct_i16model(int16_t block[64]);   /* 16-bit-stage model of avx2/fdct.rs:258-468 */
void orc_fdct_simd(int16_t block[64]);       /* AVX2 fDCT (i32 lanes) when built with AVX2, else the scalar one */
/* CPU-baseline switch: 1 = AVX2 colour conversion (Rgb/Rgba) and fDCT where the reference's `simd` feature has its
 * own (src/avx2/ycbcr.rs, src/avx2/fdct.rs); quantizer and entropy coder stay scalar, as in the reference.
 * Bit-identical output either way (tests/test_oracle.py). Not thread-safe to flip while encodes run. */
void orc_set_simd(int on);
int orc_has_simd(void);
void orc_quant_table(uint8_t kind, const uint16_t custom[64], uint8_t quality, int luma,
                     uint16_t table[64], int32_t recip[64], int32_t corr[64]); /* quantization.rs:187-283 */
int16_t orc_quantize(int16_t v, int32_t recip, int32_t corr);                 /* quantization.rs:291-307 */
uint8_t orc_get_num_bits(int16_t v);                                          /* encoder.rs:1244-1257 */
void orc_get_code(int16_t v, uint8_t *size, uint16_t *bits);                  /* writer.rs:455-470 */
/* HuffmanTable::new_optimized, huffman.rs:99-221. Returns number of values. */
int orc_huffman_optimized(const uint32_t freq[257], uint8_t length[16], uint8_t values[256]);

#ifdef __cplusplus
}
#endif
#endif
