/*
 * oracle/jpeg_oracle.c -- TEST INFRASTRUCTURE ONLY (see jpeg_oracle.h).
 *
 * Plain-C restatement of vstroebel/jpeg-encoder v0.7.0's encode path. Each function cites the
 * reference file:line it follows (paths relative to the reference's root). No code here is
 * used by the product; it is the checker and the CPU baseline.
 */
#include "jpeg_oracle.h"

#include <stdlib.h>
#include <string.h>
#if defined(__AVX2__)
#include <immintrin.h>
#endif

/* CPU-baseline switch: 1 = use the AVX2 colour conversion and fDCT below where the reference's `simd`
 * feature uses its own (src/avx2/ycbcr.rs, src/avx2/fdct.rs). Same bits either way (tests compare). */
static int g_simd = 0;
void orc_set_simd(int on) { g_simd = on; }
int orc_has_simd(void) {
#if defined(__AVX2__)
    return 1;
#else
    return 0;
#endif
}

/* ------------------------------------------------------------------------------------------
 * byte sink (the reference writes into a user `W: JfifWrite`; here: a growing buffer)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint8_t *buf;
    size_t len, cap;
    int oom;
    /* JfifWriter state, src/writer.rs:108-121 */
    uint64_t bit_buffer;
    int free_bits;
} writer;

static void w_reserve(writer *w, size_t extra) {
    if (w->len + extra <= w->cap) return;
    size_t ncap = w->cap ? w->cap * 2 : 1 << 16;
    while (ncap < w->len + extra) ncap *= 2;
    uint8_t *nb = (uint8_t *)realloc(w->buf, ncap);
    if (!nb) { w->oom = 1; return; }
    w->buf = nb;
    w->cap = ncap;
}
static void w_write(writer *w, const void *data, size_t n) {
    w_reserve(w, n);
    if (w->oom) return;
    memcpy(w->buf + w->len, data, n);
    w->len += n;
}
static void w_u8(writer *w, uint8_t v) { w_write(w, &v, 1); }
static void w_u16(writer *w, uint16_t v) { uint8_t b[2] = {(uint8_t)(v >> 8), (uint8_t)v}; w_write(w, b, 2); }
static void w_marker(writer *w, uint8_t m) { uint8_t b[2] = {0xFF, m}; w_write(w, b, 2); } /* writer.rs:204-206 */

/* src/writer.rs:156-167 */
static void flush_byte_from_bit_buffer(writer *w, int free_bits) {
    uint8_t value = (uint8_t)((w->bit_buffer >> (64 - 8 - free_bits)) & 0xFF);
    w_u8(w, value);
    if (value == 0xFF) w_u8(w, 0x00);
}
/* src/writer.rs:169-184 (the 0xFF test is a speed trick; the bytes are the same) */
static void write_bit_buffer(writer *w) {
    uint64_t b = w->bit_buffer;
    if ((b & 0x8080808080808080ull & ~(b + 0x0101010101010101ull)) != 0) {
        for (int i = 0; i < 8; i++) flush_byte_from_bit_buffer(w, i * 8);
    } else {
        uint8_t be[8];
        for (int i = 0; i < 8; i++) be[i] = (uint8_t)(b >> (56 - 8 * i));
        w_write(w, be, 8);
    }
}
/* src/writer.rs:186-202 */
static void write_bits(writer *w, uint32_t value32, uint8_t size8) {
    int size = size8;
    uint64_t value = value32;
    int free_bits = w->free_bits - size;
    if (free_bits < 0) {
        w->bit_buffer = (w->bit_buffer << (size + free_bits)) | (value >> (-free_bits));
        write_bit_buffer(w);
        w->bit_buffer = value;
        w->free_bits = free_bits + 64;
    } else {
        w->free_bits = free_bits;
        w->bit_buffer = (size == 64) ? value : ((w->bit_buffer << size) | value);
    }
}
/* src/writer.rs:147-154 */
static void flush_bit_buffer(writer *w) {
    while (w->free_bits <= 64 - 8) {
        flush_byte_from_bit_buffer(w, w->free_bits);
        w->free_bits += 8;
    }
}
/* src/writer.rs:138-145 */
static void finalize_bit_buffer(writer *w) {
    write_bits(w, 0x7F, 7);
    flush_bit_buffer(w);
    w->bit_buffer = 0;
    w->free_bits = 64;
}

/* ------------------------------------------------------------------------------------------
 * ZIGZAG, src/writer.rs:64-68
 * ---------------------------------------------------------------------------------------- */
static const uint8_t ZIGZAG[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

/* ------------------------------------------------------------------------------------------
 * colour, src/image_buffer.rs:9-38
 * ---------------------------------------------------------------------------------------- */
void orc_rgb_to_ycbcr(uint8_t r8, uint8_t g8, uint8_t b8, uint8_t out[3]) {
    int32_t r = r8, g = g8, b = b8;
    int32_t y = 19595 * r + 38470 * g + 7471 * b;
    int32_t cb = -11059 * r - 21709 * g + 32768 * b + (128 << 16);
    int32_t cr = 32768 * r - 27439 * g - 5329 * b + (128 << 16);
    y = (y + 0x7FFF) >> 16;
    cb = (cb + 0x7FFF) >> 16;
    cr = (cr + 0x7FFF) >> 16;
    out[0] = (uint8_t)y;
    out[1] = (uint8_t)cb;
    out[2] = (uint8_t)cr;
}

static int bytes_per_pixel(uint8_t ct) { /* src/encoder.rs:101-111 */
    switch (ct) {
    case ORC_LUMA: return 1;
    case ORC_RGB: case ORC_BGR: case ORC_YCBCR: return 3;
    default: return 4;
    }
}
static int num_components(uint8_t ct) { /* src/encoder.rs:55-65 + adaptors' get_jpeg_color_type */
    switch (ct) {
    case ORC_LUMA: return 1;
    case ORC_RGB: case ORC_RGBA: case ORC_BGR: case ORC_BGRA: case ORC_YCBCR: return 3;
    default: return 4;
    }
}

#if defined(__AVX2__)
/* Eight RGB(A) pixels per step (role of src/avx2/ycbcr.rs:58-125, which the `simd` feature selects for
 * the Rgb / Rgba adaptors): one gather fetches the three channel bytes of each pixel, then the same
 * i32 multiply / add / shift as orc_rgb_to_ycbcr. Returns how many pixels of the row it converted
 * (the last ones are left to the scalar loop so the 4-byte gathers never read past the row). */
static size_t rgb_row_avx2(const uint8_t *line, size_t width, int bpp, uint8_t *y, uint8_t *cb, uint8_t *cr) {
    const __m256i idx = _mm256_mullo_epi32(_mm256_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7), _mm256_set1_epi32(bpp));
    const __m256i ff = _mm256_set1_epi32(0xFF), half = _mm256_set1_epi32(0x7FFF), bias = _mm256_set1_epi32((128 << 16) + 0x7FFF);
    size_t x = 0;
    for (; x + 9 <= width; x += 8) {
        const __m256i px = _mm256_i32gather_epi32((const int *)(line + x * bpp), idx, 1);
        const __m256i r = _mm256_and_si256(px, ff), g = _mm256_and_si256(_mm256_srli_epi32(px, 8), ff),
                      b = _mm256_and_si256(_mm256_srli_epi32(px, 16), ff);
        __m256i vy = _mm256_add_epi32(_mm256_add_epi32(_mm256_mullo_epi32(r, _mm256_set1_epi32(19595)), _mm256_mullo_epi32(g, _mm256_set1_epi32(38470))),
                                      _mm256_add_epi32(_mm256_mullo_epi32(b, _mm256_set1_epi32(7471)), half));
        __m256i vcb = _mm256_add_epi32(_mm256_add_epi32(_mm256_mullo_epi32(r, _mm256_set1_epi32(-11059)), _mm256_mullo_epi32(g, _mm256_set1_epi32(-21709))),
                                       _mm256_add_epi32(_mm256_slli_epi32(b, 15), bias));
        __m256i vcr = _mm256_add_epi32(_mm256_add_epi32(_mm256_slli_epi32(r, 15), _mm256_mullo_epi32(g, _mm256_set1_epi32(-27439))),
                                       _mm256_add_epi32(_mm256_mullo_epi32(b, _mm256_set1_epi32(-5329)), bias));
        vy = _mm256_srai_epi32(vy, 16);
        vcb = _mm256_srai_epi32(vcb, 16);
        vcr = _mm256_srai_epi32(vcr, 16);
        /* 8 x i32 (0..255) -> 8 bytes */
        const __m128i y16 = _mm_packs_epi32(_mm256_castsi256_si128(vy), _mm256_extracti128_si256(vy, 1));
        const __m128i cb16 = _mm_packs_epi32(_mm256_castsi256_si128(vcb), _mm256_extracti128_si256(vcb, 1));
        const __m128i cr16 = _mm_packs_epi32(_mm256_castsi256_si128(vcr), _mm256_extracti128_si256(vcr, 1));
        _mm_storel_epi64((__m128i *)(y + x), _mm_packus_epi16(y16, y16));
        _mm_storel_epi64((__m128i *)(cb + x), _mm_packus_epi16(cb16, cb16));
        _mm_storel_epi64((__m128i *)(cr + x), _mm_packus_epi16(cr16, cr16));
    }
    return x;
}
#endif

/* ImageBuffer::fill_buffers for the nine adaptors, src/image_buffer.rs:100-313.
 * Appends `width` samples of image row y to each used plane at dst[c] and advances nothing;
 * caller owns the cursor. */
static void fill_row(const orc_params *p, const uint8_t *data, uint16_t y, uint8_t *dst[4]) {
    size_t width = p->width;
    int bpp = bytes_per_pixel(p->color_type);
    const uint8_t *line = data + (size_t)y * width * bpp; /* get_line, :125-133 */
    uint8_t t[3];
    size_t x0 = 0;
#if defined(__AVX2__)
    if (g_simd && (p->color_type == ORC_RGB || p->color_type == ORC_RGBA)) x0 = rgb_row_avx2(line, width, bpp, dst[0], dst[1], dst[2]);
#endif
    for (size_t x = x0; x < width; x++) {
        const uint8_t *px = line + x * bpp;
        switch (p->color_type) {
        case ORC_LUMA: dst[0][x] = px[0]; break;                                   /* :115-121 */
        case ORC_RGB: case ORC_RGBA:                                               /* :201-202 */
            orc_rgb_to_ycbcr(px[0], px[1], px[2], t);
            dst[0][x] = t[0]; dst[1][x] = t[1]; dst[2][x] = t[2];
            break;
        case ORC_BGR: case ORC_BGRA:                                               /* :203-204 */
            orc_rgb_to_ycbcr(px[2], px[1], px[0], t);
            dst[0][x] = t[0]; dst[1][x] = t[1]; dst[2][x] = t[2];
            break;
        case ORC_YCBCR: dst[0][x] = px[0]; dst[1][x] = px[1]; dst[2][x] = px[2]; break; /* :221-229 */
        case ORC_CMYK:                                                             /* :247-256 */
            dst[0][x] = 255 - px[0]; dst[1][x] = 255 - px[1];
            dst[2][x] = 255 - px[2]; dst[3][x] = 255 - px[3];
            break;
        case ORC_CMYK_AS_YCCK:                                                     /* :274-285, :35-38 */
            orc_rgb_to_ycbcr(px[0], px[1], px[2], t);
            dst[0][x] = t[0]; dst[1][x] = t[1]; dst[2][x] = t[2]; dst[3][x] = 255 - px[3];
            break;
        default: /* ORC_YCCK :303-312 */
            dst[0][x] = px[0]; dst[1][x] = px[1]; dst[2][x] = px[2]; dst[3][x] = px[3];
            break;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * fDCT, src/fdct.rs:76-238
 * ---------------------------------------------------------------------------------------- */
#define CONST_BITS 13
#define PASS1_BITS 2
#define FIX_0_298631336 2446
#define FIX_0_390180644 3196
#define FIX_0_541196100 4433
#define FIX_0_765366865 6270
#define FIX_0_899976223 7373
#define FIX_1_175875602 9633
#define FIX_1_501321110 12299
#define FIX_1_847759065 15137
#define FIX_1_961570560 16069
#define FIX_2_053119869 16819
#define FIX_2_562915447 20995
#define FIX_3_072711026 25172

static inline int32_t descale(int32_t x, int n) { return (x + (1 << (n - 1))) >> n; } /* fdct.rs:94-98 */

static void fdct_scalar(int16_t data[64]) {
    int32_t data2[64];
    for (int y = 0; y < 8; y++) { /* pass 1: rows, fdct.rs:116-171 */
        int o = y * 8;
        int32_t tmp0 = data[o + 0] + data[o + 7], tmp7 = data[o + 0] - data[o + 7];
        int32_t tmp1 = data[o + 1] + data[o + 6], tmp6 = data[o + 1] - data[o + 6];
        int32_t tmp2 = data[o + 2] + data[o + 5], tmp5 = data[o + 2] - data[o + 5];
        int32_t tmp3 = data[o + 3] + data[o + 4], tmp4 = data[o + 3] - data[o + 4];
        int32_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        data2[o + 0] = (tmp10 + tmp11) * (1 << PASS1_BITS);
        data2[o + 4] = (tmp10 - tmp11) * (1 << PASS1_BITS);
        int32_t z1 = (tmp12 + tmp13) * FIX_0_541196100;
        data2[o + 2] = descale(z1 + tmp13 * FIX_0_765366865, CONST_BITS - PASS1_BITS);
        data2[o + 6] = descale(z1 + tmp12 * -FIX_1_847759065, CONST_BITS - PASS1_BITS);
        z1 = tmp4 + tmp7;
        int32_t z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
        int32_t z5 = (z3 + z4) * FIX_1_175875602;
        tmp4 *= FIX_0_298631336; tmp5 *= FIX_2_053119869; tmp6 *= FIX_3_072711026; tmp7 *= FIX_1_501321110;
        z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
        z3 += z5; z4 += z5;
        data2[o + 7] = descale(tmp4 + z1 + z3, CONST_BITS - PASS1_BITS);
        data2[o + 5] = descale(tmp5 + z2 + z4, CONST_BITS - PASS1_BITS);
        data2[o + 3] = descale(tmp6 + z2 + z3, CONST_BITS - PASS1_BITS);
        data2[o + 1] = descale(tmp7 + z1 + z4, CONST_BITS - PASS1_BITS);
    }
    for (int x = 0; x < 8; x++) { /* pass 2: columns, fdct.rs:178-237 */
        int32_t tmp0 = data2[0 + x] + data2[56 + x], tmp7 = data2[0 + x] - data2[56 + x];
        int32_t tmp1 = data2[8 + x] + data2[48 + x], tmp6 = data2[8 + x] - data2[48 + x];
        int32_t tmp2 = data2[16 + x] + data2[40 + x], tmp5 = data2[16 + x] - data2[40 + x];
        int32_t tmp3 = data2[24 + x] + data2[32 + x], tmp4 = data2[24 + x] - data2[32 + x];
        int32_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        data[0 + x] = (int16_t)descale(tmp10 + tmp11, PASS1_BITS);
        data[32 + x] = (int16_t)descale(tmp10 - tmp11, PASS1_BITS);
        int32_t z1 = (tmp12 + tmp13) * FIX_0_541196100;
        data[16 + x] = (int16_t)descale(z1 + tmp13 * FIX_0_765366865, CONST_BITS + PASS1_BITS);
        data[48 + x] = (int16_t)descale(z1 + tmp12 * -FIX_1_847759065, CONST_BITS + PASS1_BITS);
        z1 = tmp4 + tmp7;
        int32_t z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
        int32_t z5 = (z3 + z4) * FIX_1_175875602;
        tmp4 *= FIX_0_298631336; tmp5 *= FIX_2_053119869; tmp6 *= FIX_3_072711026; tmp7 *= FIX_1_501321110;
        z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
        z3 += z5; z4 += z5;
        data[56 + x] = (int16_t)descale(tmp4 + z1 + z3, CONST_BITS + PASS1_BITS);
        data[40 + x] = (int16_t)descale(tmp5 + z2 + z4, CONST_BITS + PASS1_BITS);
        data[24 + x] = (int16_t)descale(tmp6 + z2 + z3, CONST_BITS + PASS1_BITS);
        data[8 + x] = (int16_t)descale(tmp7 + z1 + z4, CONST_BITS + PASS1_BITS);
    }
}

#if defined(__AVX2__)
/* The same LL&M flow on eight lanes of i32 (role of src/avx2/fdct.rs, which works in 16-bit lanes; i32 lanes
 * make bit-equality with the scalar code a matter of construction). v[k] holds element k of eight independent
 * 8-point transforms. pass 1: outputs scaled by 2^PASS1_BITS; pass 2: that scale removed again. */
static inline __m256i vdescale(__m256i x, int n) { return _mm256_srai_epi32(_mm256_add_epi32(x, _mm256_set1_epi32(1 << (n - 1))), n); }
#define VMUL(a, c) _mm256_mullo_epi32((a), _mm256_set1_epi32(c))
static void dct_pass_avx2(__m256i v[8], int first) {
    const int sh = first ? CONST_BITS - PASS1_BITS : CONST_BITS + PASS1_BITS;
    const __m256i tmp0 = _mm256_add_epi32(v[0], v[7]), tmp7 = _mm256_sub_epi32(v[0], v[7]);
    const __m256i tmp1 = _mm256_add_epi32(v[1], v[6]), tmp6 = _mm256_sub_epi32(v[1], v[6]);
    const __m256i tmp2 = _mm256_add_epi32(v[2], v[5]), tmp5 = _mm256_sub_epi32(v[2], v[5]);
    const __m256i tmp3 = _mm256_add_epi32(v[3], v[4]), tmp4 = _mm256_sub_epi32(v[3], v[4]);
    const __m256i tmp10 = _mm256_add_epi32(tmp0, tmp3), tmp13 = _mm256_sub_epi32(tmp0, tmp3);
    const __m256i tmp11 = _mm256_add_epi32(tmp1, tmp2), tmp12 = _mm256_sub_epi32(tmp1, tmp2);
    if (first) {
        v[0] = _mm256_slli_epi32(_mm256_add_epi32(tmp10, tmp11), PASS1_BITS);
        v[4] = _mm256_slli_epi32(_mm256_sub_epi32(tmp10, tmp11), PASS1_BITS);
    } else {
        v[0] = vdescale(_mm256_add_epi32(tmp10, tmp11), PASS1_BITS);
        v[4] = vdescale(_mm256_sub_epi32(tmp10, tmp11), PASS1_BITS);
    }
    __m256i z1 = VMUL(_mm256_add_epi32(tmp12, tmp13), FIX_0_541196100);
    v[2] = vdescale(_mm256_add_epi32(z1, VMUL(tmp13, FIX_0_765366865)), sh);
    v[6] = vdescale(_mm256_add_epi32(z1, VMUL(tmp12, -FIX_1_847759065)), sh);
    z1 = _mm256_add_epi32(tmp4, tmp7);
    __m256i z2 = _mm256_add_epi32(tmp5, tmp6), z3 = _mm256_add_epi32(tmp4, tmp6), z4 = _mm256_add_epi32(tmp5, tmp7);
    const __m256i z5 = VMUL(_mm256_add_epi32(z3, z4), FIX_1_175875602);
    const __m256i t4 = VMUL(tmp4, FIX_0_298631336), t5 = VMUL(tmp5, FIX_2_053119869), t6 = VMUL(tmp6, FIX_3_072711026),
                  t7 = VMUL(tmp7, FIX_1_501321110);
    z1 = VMUL(z1, -FIX_0_899976223);
    z2 = VMUL(z2, -FIX_2_562915447);
    z3 = _mm256_add_epi32(VMUL(z3, -FIX_1_961570560), z5);
    z4 = _mm256_add_epi32(VMUL(z4, -FIX_0_390180644), z5);
    v[7] = vdescale(_mm256_add_epi32(_mm256_add_epi32(t4, z1), z3), sh);
    v[5] = vdescale(_mm256_add_epi32(_mm256_add_epi32(t5, z2), z4), sh);
    v[3] = vdescale(_mm256_add_epi32(_mm256_add_epi32(t6, z2), z3), sh);
    v[1] = vdescale(_mm256_add_epi32(_mm256_add_epi32(t7, z1), z4), sh);
}
static void transpose8_avx2(__m256i r[8]) {
    const __m256i a0 = _mm256_unpacklo_epi32(r[0], r[1]), a1 = _mm256_unpackhi_epi32(r[0], r[1]);
    const __m256i a2 = _mm256_unpacklo_epi32(r[2], r[3]), a3 = _mm256_unpackhi_epi32(r[2], r[3]);
    const __m256i a4 = _mm256_unpacklo_epi32(r[4], r[5]), a5 = _mm256_unpackhi_epi32(r[4], r[5]);
    const __m256i a6 = _mm256_unpacklo_epi32(r[6], r[7]), a7 = _mm256_unpackhi_epi32(r[6], r[7]);
    const __m256i b0 = _mm256_unpacklo_epi64(a0, a2), b1 = _mm256_unpackhi_epi64(a0, a2);
    const __m256i b2 = _mm256_unpacklo_epi64(a1, a3), b3 = _mm256_unpackhi_epi64(a1, a3);
    const __m256i b4 = _mm256_unpacklo_epi64(a4, a6), b5 = _mm256_unpackhi_epi64(a4, a6);
    const __m256i b6 = _mm256_unpacklo_epi64(a5, a7), b7 = _mm256_unpackhi_epi64(a5, a7);
    r[0] = _mm256_permute2x128_si256(b0, b4, 0x20);
    r[1] = _mm256_permute2x128_si256(b1, b5, 0x20);
    r[2] = _mm256_permute2x128_si256(b2, b6, 0x20);
    r[3] = _mm256_permute2x128_si256(b3, b7, 0x20);
    r[4] = _mm256_permute2x128_si256(b0, b4, 0x31);
    r[5] = _mm256_permute2x128_si256(b1, b5, 0x31);
    r[6] = _mm256_permute2x128_si256(b2, b6, 0x31);
    r[7] = _mm256_permute2x128_si256(b3, b7, 0x31);
}
static void fdct_avx2(int16_t data[64]) {
    __m256i v[8];
    for (int i = 0; i < 8; i++) v[i] = _mm256_cvtepi16_epi32(_mm_loadu_si128((const __m128i *)(data + 8 * i)));
    transpose8_avx2(v);   /* v[k] = column k: element k of every row */
    dct_pass_avx2(v, 1);  /* rows, fdct.rs:116-171 */
    transpose8_avx2(v);   /* v[k] = row k of the intermediate */
    dct_pass_avx2(v, 0);  /* columns, fdct.rs:178-237 */
    for (int i = 0; i < 8; i++) /* results fit i16 (|v| <= 8192), so the saturating pack is the `as i16` of the scalar code */
        _mm_storeu_si128((__m128i *)(data + 8 * i), _mm_packs_epi32(_mm256_castsi256_si128(v[i]), _mm256_extracti128_si256(v[i], 1)));
}
void orc_fdct_simd(int16_t data[64]) { fdct_avx2(data); }
#else
void orc_fdct_simd(int16_t data[64]) { fdct_scalar(data); }
#endif

void orc_fdct(int16_t data[64]) {
#if defined(__AVX2__)
    if (g_simd) {
        fdct_avx2(data);
        return;
    }
#endif
    fdct_scalar(data);
}

/* 16-bit-stage model of the AVX2 backend (src/avx2/fdct.rs:258-423): every `_epi16` add/sub/shift
 * wraps to i16, `madd_epi16` sums two i16*i16 products in i32, `packs_epi32` saturates to i16.
 * Used by tests to confirm that the `simd` feature produces the same coefficients as the scalar
 * path for 8-bit samples (the reference holds no such test). */
static inline int16_t wrap16(int32_t v) { return (int16_t)(uint16_t)(uint32_t)v; }
static inline int16_t sat16(int32_t v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : (int16_t)v); }
static void dct1d_i16model(const int16_t in[8], int16_t out[8], int first_pass) {
    int16_t tmp0 = wrap16(in[0] + in[7]), tmp7 = wrap16(in[0] - in[7]);
    int16_t tmp1 = wrap16(in[1] + in[6]), tmp6 = wrap16(in[1] - in[6]);
    int16_t tmp2 = wrap16(in[2] + in[5]), tmp5 = wrap16(in[2] - in[5]);
    int16_t tmp3 = wrap16(in[3] + in[4]), tmp4 = wrap16(in[3] - in[4]);
    int16_t tmp10 = wrap16(tmp0 + tmp3), tmp13 = wrap16(tmp0 - tmp3);
    int16_t tmp11 = wrap16(tmp1 + tmp2), tmp12 = wrap16(tmp1 - tmp2);
    int16_t s = wrap16(tmp10 + tmp11), d = wrap16(tmp10 - tmp11);
    if (first_pass) {
        out[0] = wrap16((int32_t)s * 4);
        out[4] = wrap16((int32_t)d * 4);
    } else {
        out[0] = (int16_t)(wrap16(s + 2) >> PASS1_BITS);
        out[4] = (int16_t)(wrap16(d + 2) >> PASS1_BITS);
    }
    int n = first_pass ? CONST_BITS - PASS1_BITS : CONST_BITS + PASS1_BITS;
    int32_t rnd = 1 << (n - 1);
    /* avx2/fdct.rs:296-329 pre-combined constants */
    out[2] = sat16((tmp13 * (int32_t)(int16_t)(FIX_0_541196100 + FIX_0_765366865) + tmp12 * (int32_t)FIX_0_541196100 + rnd) >> n);
    out[6] = sat16((tmp13 * (int32_t)FIX_0_541196100 + tmp12 * (int32_t)(int16_t)(FIX_0_541196100 - FIX_1_847759065) + rnd) >> n);
    int16_t z3 = wrap16(tmp4 + tmp6), z4 = wrap16(tmp5 + tmp7);
    int32_t z3m = z3 * (int32_t)(int16_t)(FIX_1_175875602 - FIX_1_961570560) + z4 * (int32_t)FIX_1_175875602;
    int32_t z4m = z3 * (int32_t)FIX_1_175875602 + z4 * (int32_t)(int16_t)(FIX_1_175875602 - FIX_0_390180644);
    int32_t t4 = tmp4 * (int32_t)(int16_t)(FIX_0_298631336 - FIX_0_899976223) + tmp7 * (int32_t)(int16_t)(-FIX_0_899976223);
    int32_t t5 = tmp5 * (int32_t)(int16_t)(FIX_2_053119869 - FIX_2_562915447) + tmp6 * (int32_t)(int16_t)(-FIX_2_562915447);
    int32_t t6 = tmp5 * (int32_t)(int16_t)(-FIX_2_562915447) + tmp6 * (int32_t)(int16_t)(FIX_3_072711026 - FIX_2_562915447);
    int32_t t7 = tmp4 * (int32_t)(int16_t)(-FIX_0_899976223) + tmp7 * (int32_t)(int16_t)(FIX_1_501321110 - FIX_0_899976223);
    out[7] = sat16((t4 + z3m + rnd) >> n);
    out[5] = sat16((t5 + z4m + rnd) >> n);
    out[3] = sat16((t6 + z3m + rnd) >> n);
    out[1] = sat16((t7 + z4m + rnd) >> n);
}
void orc_fdct_i16model(int16_t data[64]) {
    int16_t tmp[64], in[8], out[8];
    for (int y = 0; y < 8; y++) {
        dct1d_i16model(data + y * 8, out, 1);
        memcpy(tmp + y * 8, out, sizeof(out));
    }
    for (int x = 0; x < 8; x++) {
        for (int y = 0; y < 8; y++) in[y] = tmp[y * 8 + x];
        dct1d_i16model(in, out, 0);
        for (int y = 0; y < 8; y++) data[y * 8 + x] = out[y];
    }
}

/* ------------------------------------------------------------------------------------------
 * quantization, src/quantization.rs
 * ---------------------------------------------------------------------------------------- */
static const uint16_t LUMA_TABLES[9][64] = { /* quantization.rs:62-121 */
    {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99},
    {16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16},
    {12, 17, 20, 21, 30, 34, 56, 63, 18, 20, 20, 26, 28, 51, 61, 55, 19, 20, 21, 26, 33, 58, 69, 55, 26, 26, 26, 30, 46, 87, 86, 66, 31, 33, 36, 40, 46, 96, 100, 73, 40, 35, 46, 62, 81, 100, 111, 91, 46, 66, 76, 86, 102, 121, 120, 101, 68, 90, 90, 96, 113, 102, 105, 103},
    {9, 10, 12, 14, 27, 32, 51, 62, 11, 12, 14, 19, 27, 44, 59, 73, 12, 14, 18, 25, 42, 59, 79, 78, 17, 18, 25, 42, 61, 92, 87, 92, 23, 28, 42, 75, 79, 112, 112, 99, 40, 42, 59, 84, 88, 124, 132, 111, 42, 64, 78, 95, 105, 126, 125, 99, 70, 75, 100, 102, 116, 100, 107, 98},
    {16, 16, 16, 18, 25, 37, 56, 85, 16, 17, 20, 27, 34, 40, 53, 75, 16, 20, 24, 31, 43, 62, 91, 135, 18, 27, 31, 40, 53, 74, 106, 156, 25, 34, 43, 53, 69, 94, 131, 189, 37, 40, 62, 74, 94, 124, 169, 238, 56, 53, 91, 106, 131, 169, 226, 311, 85, 75, 135, 156, 189, 238, 311, 418},
    {10, 12, 14, 19, 26, 38, 57, 86, 12, 18, 21, 28, 35, 41, 54, 76, 14, 21, 25, 32, 44, 63, 92, 136, 19, 28, 32, 41, 54, 75, 107, 157, 26, 35, 44, 54, 70, 95, 132, 190, 38, 41, 63, 75, 95, 125, 170, 239, 57, 54, 92, 107, 132, 170, 227, 312, 86, 76, 136, 157, 190, 239, 312, 419},
    {7, 8, 10, 14, 23, 44, 95, 241, 8, 8, 11, 15, 25, 47, 102, 255, 10, 11, 13, 19, 31, 58, 127, 255, 14, 15, 19, 27, 44, 83, 181, 255, 23, 25, 31, 44, 72, 136, 255, 255, 44, 47, 58, 83, 136, 255, 255, 255, 95, 102, 127, 181, 255, 255, 255, 255, 241, 255, 255, 255, 255, 255, 255, 255},
    {15, 11, 11, 12, 15, 19, 25, 32, 11, 13, 10, 10, 12, 15, 19, 24, 11, 10, 14, 14, 16, 18, 22, 27, 12, 10, 14, 18, 21, 24, 28, 33, 15, 12, 16, 21, 26, 31, 36, 42, 19, 15, 18, 24, 31, 38, 45, 53, 25, 19, 22, 28, 36, 45, 55, 65, 32, 24, 27, 33, 42, 53, 65, 77},
    {14, 10, 11, 14, 19, 25, 34, 45, 10, 11, 11, 12, 15, 20, 26, 33, 11, 11, 15, 18, 21, 25, 31, 38, 14, 12, 18, 24, 28, 33, 39, 47, 19, 15, 21, 28, 36, 43, 51, 59, 25, 20, 25, 33, 43, 54, 64, 74, 34, 26, 31, 39, 51, 64, 77, 91, 45, 33, 38, 47, 59, 74, 91, 108},
};
static const uint16_t CHROMA_TABLES[9][64] = { /* quantization.rs:124-183 */
    {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99},
    {16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16},
    {8, 12, 15, 15, 86, 96, 96, 98, 13, 13, 15, 26, 90, 96, 99, 98, 12, 15, 18, 96, 99, 99, 99, 99, 17, 16, 90, 96, 99, 99, 99, 99, 96, 96, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99},
    {9, 10, 17, 19, 62, 89, 91, 97, 12, 13, 18, 29, 84, 91, 88, 98, 14, 19, 29, 93, 95, 95, 98, 97, 20, 26, 84, 88, 95, 95, 98, 94, 26, 86, 91, 93, 97, 99, 98, 99, 99, 100, 98, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 97, 97, 99, 99, 99, 99, 97, 99},
    {16, 16, 16, 18, 25, 37, 56, 85, 16, 17, 20, 27, 34, 40, 53, 75, 16, 20, 24, 31, 43, 62, 91, 135, 18, 27, 31, 40, 53, 74, 106, 156, 25, 34, 43, 53, 69, 94, 131, 189, 37, 40, 62, 74, 94, 124, 169, 238, 56, 53, 91, 106, 131, 169, 226, 311, 85, 75, 135, 156, 189, 238, 311, 418},
    {10, 12, 14, 19, 26, 38, 57, 86, 12, 18, 21, 28, 35, 41, 54, 76, 14, 21, 25, 32, 44, 63, 92, 136, 19, 28, 32, 41, 54, 75, 107, 157, 26, 35, 44, 54, 70, 95, 132, 190, 38, 41, 63, 75, 95, 125, 170, 239, 57, 54, 92, 107, 132, 170, 227, 312, 86, 76, 136, 157, 190, 239, 312, 419},
    {7, 8, 10, 14, 23, 44, 95, 241, 8, 8, 11, 15, 25, 47, 102, 255, 10, 11, 13, 19, 31, 58, 127, 255, 14, 15, 19, 27, 44, 83, 181, 255, 23, 25, 31, 44, 72, 136, 255, 255, 44, 47, 58, 83, 136, 255, 255, 255, 95, 102, 127, 181, 255, 255, 255, 255, 241, 255, 255, 255, 255, 255, 255, 255},
    {15, 11, 11, 12, 15, 19, 25, 32, 11, 13, 10, 10, 12, 15, 19, 24, 11, 10, 14, 14, 16, 18, 22, 27, 12, 10, 14, 18, 21, 24, 28, 33, 15, 12, 16, 21, 26, 31, 36, 42, 19, 15, 18, 24, 31, 38, 45, 53, 25, 19, 22, 28, 36, 45, 55, 65, 32, 24, 27, 33, 42, 53, 65, 77},
    {14, 10, 11, 14, 19, 25, 34, 45, 10, 11, 11, 12, 15, 20, 26, 33, 11, 11, 15, 18, 21, 25, 31, 38, 14, 12, 18, 24, 28, 33, 39, 47, 19, 15, 21, 28, 36, 43, 51, 59, 25, 20, 25, 33, 43, 54, 64, 74, 34, 26, 31, 39, 51, 64, 77, 91, 45, 33, 38, 47, 59, 74, 91, 108},
};

/* quantization.rs:187-207 */
static void compute_reciprocal(uint32_t divisor, int32_t *recip, int32_t *corr) {
    if (divisor <= 1) { *recip = 1; *corr = 0; return; }
    uint32_t reciprocals = (1u << 15) / divisor;
    uint32_t fractional = (1u << 15) % divisor;
    uint32_t correction = divisor / 2;
    if (fractional != 0) {
        if (fractional <= correction) correction += 1;
        else reciprocals += 1;
    }
    *recip = (int32_t)reciprocals;
    *corr = (int32_t)correction;
}

/* QuantizationTable::new_with_quality, quantization.rs:216-283. table[] holds the value << 3. */
void orc_quant_table(uint8_t kind, const uint16_t custom[64], uint8_t quality, int luma,
                     uint16_t table[64], int32_t recip[64], int32_t corr[64]) {
    if (kind >= 9) { /* get_user_table :250-259 */
        for (int i = 0; i < 64; i++) {
            uint16_t v = custom[i];
            if (v < 1) v = 1;
            if (v > (2 << 10)) v = 2 << 10;
            table[i] = (uint16_t)(v << 3);
        }
    } else { /* get_with_quality :261-283 */
        const uint16_t *base = luma ? LUMA_TABLES[kind] : CHROMA_TABLES[kind];
        uint32_t q = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
        uint32_t scale = q < 50 ? 5000 / q : 200 - q * 2;
        for (int i = 0; i < 64; i++) {
            uint32_t v = ((uint32_t)base[i] * scale + 50) / 100;
            if (v < 1) v = 1;
            if (v > 255) v = 255;
            table[i] = (uint16_t)(v << 3);
        }
    }
    for (int i = 0; i < 64; i++) compute_reciprocal(table[i], &recip[i], &corr[i]);
}

/* QuantizationTable::quantize, quantization.rs:291-307 */
int16_t orc_quantize(int16_t in_value, int32_t reciprocal, int32_t corrections) {
    int32_t value = in_value;
    int32_t abs_value = value < 0 ? -value : value;
    int32_t product = (abs_value + corrections) * reciprocal;
    product >>= 15;
    if (value != abs_value) product *= -1;
    return (int16_t)product;
}

typedef struct {
    uint16_t table[64];
    int32_t recip[64], corr[64];
} qtable;

/* Operations::quantize_block, src/encoder.rs:1265-1271 (output is in zig-zag order) */
static void quantize_block(const int16_t block[64], int16_t q_block[64], const qtable *t) {
    for (int i = 0; i < 64; i++) {
        int z = ZIGZAG[i] & 0x3f;
        q_block[i] = orc_quantize(block[z], t->recip[z], t->corr[z]);
    }
}

/* ------------------------------------------------------------------------------------------
 * Huffman tables, src/huffman.rs
 * ---------------------------------------------------------------------------------------- */
static const uint8_t LUMA_DC_LEN[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t DC_VALUES[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t CHROMA_DC_LEN[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t LUMA_AC_LEN[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7D};
static const uint8_t LUMA_AC_VALUES[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xA1, 0x08, 0x23, 0x42, 0xB1, 0xC1, 0x15, 0x52, 0xD1, 0xF0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0A, 0x16, 0x17, 0x18, 0x19, 0x1A, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2A, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A, 0xA2, 0xA3,
    0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA, 0xC2, 0xC3,
    0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA, 0xE1, 0xE2,
    0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF1, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9, 0xFA};
static const uint8_t CHROMA_AC_LEN[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t CHROMA_AC_VALUES[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xA1, 0xB1, 0xC1, 0x09, 0x23, 0x33, 0x52, 0xF0, 0x15, 0x62, 0x72, 0xD1,
    0x0A, 0x16, 0x24, 0x34, 0xE1, 0x25, 0xF1, 0x17, 0x18, 0x19, 0x1A, 0x26, 0x27, 0x28, 0x29, 0x2A, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A,
    0xA2, 0xA3, 0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA,
    0xC2, 0xC3, 0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA,
    0xE2, 0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9, 0xFA};

typedef struct {
    uint8_t lut_size[256];  /* lookup_table.0 */
    uint16_t lut_code[256]; /* lookup_table.1 */
    uint8_t length[16];
    uint8_t values[256];
    int n_values;
} huff_table;

/* create_sizes / create_codes / create_lookup_table, huffman.rs:240-288 */
static void huff_build_lookup(huff_table *t) {
    uint8_t sizes[256];
    uint16_t codes[256];
    memset(sizes, 0, sizeof(sizes));
    memset(codes, 0, sizeof(codes));
    int k = 0;
    for (int i = 0; i < 16; i++)
        for (int j = 0; j < t->length[i]; j++) sizes[k++] = (uint8_t)(i + 1);
    uint32_t current_code = 0;
    uint8_t current_size = sizes[0];
    for (int i = 0; i < 256 && sizes[i] != 0; i++) {
        if (current_size != sizes[i]) {
            current_code <<= (sizes[i] - current_size);
            current_size = sizes[i];
        }
        codes[i] = (uint16_t)current_code;
        current_code += 1;
    }
    memset(t->lut_size, 0, sizeof(t->lut_size));
    memset(t->lut_code, 0, sizeof(t->lut_code));
    for (int i = 0; i < t->n_values; i++) {
        t->lut_size[t->values[i]] = sizes[i];
        t->lut_code[t->values[i]] = codes[i];
    }
}
static void huff_new(huff_table *t, const uint8_t length[16], const uint8_t *values, int n) { /* :73-79 */
    memcpy(t->length, length, 16);
    memset(t->values, 0, sizeof(t->values));
    memcpy(t->values, values, (size_t)n);
    t->n_values = n;
    huff_build_lookup(t);
}

/* HuffmanTable::new_optimized, huffman.rs:99-221 */
int orc_huffman_optimized(const uint32_t freq_in[257], uint8_t length[16], uint8_t values[256]) {
    uint32_t freq[257];
    int32_t others[257];
    uint32_t codesize[257];
    memcpy(freq, freq_in, sizeof(freq));
    for (int i = 0; i < 257; i++) { others[i] = -1; codesize[i] = 0; }
    for (;;) { /* Figure K.1 */
        int v1 = -1, v2 = -1;
        uint32_t v1_min = 0xFFFFFFFFu, v2_min = 0xFFFFFFFFu;
        for (int i = 0; i < 257; i++)
            if (freq[i] > 0 && freq[i] <= v1_min) { v1_min = freq[i]; v1 = i; }
        if (v1 < 0) break;
        for (int i = 0; i < 257; i++)
            if (freq[i] > 0 && freq[i] <= v2_min && i != v1) { v2_min = freq[i]; v2 = i; }
        if (v2 < 0) break;
        freq[v1] += freq[v2];
        freq[v2] = 0;
        codesize[v1] += 1;
        while (others[v1] >= 0) { v1 = others[v1]; codesize[v1] += 1; }
        others[v1] = v2;
        codesize[v2] += 1;
        while (others[v2] >= 0) { v2 = others[v2]; codesize[v2] += 1; }
    }
    uint8_t bits[33]; /* Figure K.2 */
    memset(bits, 0, sizeof(bits));
    for (int i = 0; i < 257; i++)
        if (codesize[i] > 0) {
            if (codesize[i] > 32) return -1; /* the reference would panic (index out of bounds) */
            bits[codesize[i]] += 1;
        }
    int i = 32; /* Figure K.3 */
    while (i > 16) {
        while (bits[i] > 0) {
            int j = i - 2;
            while (bits[j] == 0) j -= 1;
            bits[i] -= 2;
            bits[i - 1] += 1;
            bits[j + 1] += 2;
            bits[j] -= 1;
        }
        i -= 1;
    }
    while (bits[i] == 0) i -= 1;
    bits[i] -= 1;
    int k = 0; /* Figure K.4 */
    for (uint32_t sz = 1; sz <= 32; sz++)
        for (int j = 0; j <= 255; j++)
            if (codesize[j] == sz) values[k++] = (uint8_t)j;
    for (int l = 0; l < 16; l++) length[l] = bits[l + 1];
    return k;
}

/* ------------------------------------------------------------------------------------------
 * entropy coding, src/writer.rs:308-388, 455-470; src/encoder.rs:1244-1257
 * ---------------------------------------------------------------------------------------- */
void orc_get_code(int16_t value, uint8_t *size, uint16_t *bits) { /* writer.rs:455-470 */
    int16_t temp = (int16_t)(value - (value < 0 ? 1 : 0));
    uint16_t temp2 = (uint16_t)(value < 0 ? -value : value);
    uint16_t x = (uint16_t)((uint16_t)(temp2 << 1) | 1);
    int lz = 0;
    while (!(x & 0x8000)) { x <<= 1; lz++; }
    int num_bits = 15 - lz;
    uint16_t mask = (uint16_t)((1u << num_bits) - 1);
    *size = (uint8_t)num_bits;
    *bits = (uint16_t)((uint16_t)temp & mask);
}
uint8_t orc_get_num_bits(int16_t value16) { /* encoder.rs:1244-1257 */
    int value = value16;
    if (value < 0) value = -value;
    uint8_t n = 0;
    while (value > 0) { n += 1; value >>= 1; }
    return n;
}
static void huffman_encode(writer *w, uint8_t val, const huff_table *t) { /* writer.rs:308-312 */
    write_bits(w, t->lut_code[val], t->lut_size[val]);
}
static void huffman_encode_value(writer *w, uint8_t size, uint8_t symbol, uint16_t value, const huff_table *t) {
    uint8_t num_bits = t->lut_size[symbol]; /* writer.rs:314-329; size 0 == missing code (quirk Q18) */
    uint32_t temp = value;
    temp |= (uint32_t)t->lut_code[symbol] << size;
    write_bits(w, temp, (uint8_t)(size + num_bits));
}
static void write_dc(writer *w, int16_t value, int16_t prev_dc, const huff_table *dc) { /* writer.rs:342-354 */
    int16_t diff = (int16_t)(value - prev_dc);
    uint8_t size;
    uint16_t bits;
    orc_get_code(diff, &size, &bits);
    huffman_encode_value(w, size, size, bits, dc);
}
static void write_ac_block(writer *w, const int16_t *block, int start, int end, const huff_table *ac) {
    int zero_run = 0; /* writer.rs:356-388 */
    for (int i = start; i < end; i++) {
        int16_t value = block[i];
        if (value == 0) {
            zero_run += 1;
        } else {
            while (zero_run > 15) {
                huffman_encode(w, 0xF0, ac);
                zero_run -= 16;
            }
            uint8_t size;
            uint16_t bits;
            orc_get_code(value, &size, &bits);
            uint8_t symbol = (uint8_t)((zero_run << 4) | size);
            huffman_encode_value(w, size, symbol, bits, ac);
            zero_run = 0;
        }
    }
    if (zero_run > 0) huffman_encode(w, 0x00, ac);
}
static void write_block(writer *w, const int16_t *block, int16_t prev_dc, const huff_table *dc, const huff_table *ac) {
    write_dc(w, block[0], prev_dc, dc); /* writer.rs:331-340 */
    write_ac_block(w, block, 1, 64, ac);
}

/* ------------------------------------------------------------------------------------------
 * encoder state, src/encoder.rs:190-231
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint8_t id, quantization_table, dc_huffman_table, ac_huffman_table;
    uint8_t h, v;
} component;

typedef struct {
    const orc_params *p;
    const uint8_t *data;
    writer w;
    component comps[4];
    int ncomp;
    huff_table huff[2][2]; /* [table][0=dc,1=ac] */
    qtable q[2];
    int max_h, max_v;
} encoder;

static void add_component(encoder *e, uint8_t id, uint8_t dest, uint8_t h, uint8_t v) { /* :199-210 */
    component *c = &e->comps[e->ncomp++];
    c->id = id; c->quantization_table = dest; c->dc_huffman_table = dest; c->ac_huffman_table = dest;
    c->h = h; c->v = v;
}
static void init_components(encoder *e) { /* :569-619 */
    uint8_t value = e->p->sampling;
    uint8_t h = (value >> 4) & 0x07, v = value & 0xf; /* get_sampling_factors :173-176 */
    e->ncomp = 0;
    switch (num_components(e->p->color_type)) {
    case 1: add_component(e, 0, 0, 1, 1); break;
    case 3:
        add_component(e, 0, 0, h, v); add_component(e, 1, 1, 1, 1); add_component(e, 2, 1, 1, 1);
        break;
    default:
        if (e->p->color_type == ORC_CMYK) {
            add_component(e, 0, 1, 1, 1); add_component(e, 1, 1, 1, 1); add_component(e, 2, 1, 1, 1);
            add_component(e, 3, 0, h, v);
        } else {
            add_component(e, 0, 0, h, v); add_component(e, 1, 1, 1, 1); add_component(e, 2, 1, 1, 1);
            add_component(e, 3, 0, h, v);
        }
    }
    e->max_h = 1; e->max_v = 1; /* get_max_sampling_size :621-631 */
    for (int i = 0; i < e->ncomp; i++) {
        if (e->comps[i].h > e->max_h) e->max_h = e->comps[i].h;
        if (e->comps[i].v > e->max_v) e->max_v = e->comps[i].v;
    }
}
static int supports_interleaved(uint8_t sampling) { /* :178-187 */
    uint8_t s = sampling & 0x7f;
    return s == 0x11 || s == 0x21 || s == 0x12 || s == 0x22;
}

/* get_block, src/encoder.rs:1222-1242 */
static void get_block(const uint8_t *data, size_t start_x, size_t start_y, size_t col_stride, size_t row_stride,
                      size_t width, int16_t block[64]) {
    for (size_t y = 0; y < 8; y++)
        for (size_t x = 0; x < 8; x++) {
            size_t ix = start_x + x * col_stride, iy = start_y + y * row_stride;
            block[y * 8 + x] = (int16_t)((int16_t)data[iy * width + ix] - 128);
        }
}

/* container segments, src/writer.rs:208-306, 390-452 */
static void write_segment(writer *w, uint8_t marker, const uint8_t *data, size_t n) {
    w_marker(w, marker);
    w_u16(w, (uint16_t)(n + 2));
    w_write(w, data, n);
}
static void write_header(writer *w, const orc_params *p) { /* writer.rs:216-239 */
    w_marker(w, 0xE0);
    w_u16(w, 16);
    w_write(w, "JFIF\0", 5);
    uint8_t ver[2] = {0x01, 0x02};
    w_write(w, ver, 2);
    w_u8(w, p->density_unit == 1 ? 0x01 : (p->density_unit == 2 ? 0x02 : 0x00));
    w_u16(w, p->density_x);
    w_u16(w, p->density_y);
    uint8_t z[2] = {0, 0};
    w_write(w, z, 2);
}
static void write_huffman_segment(writer *w, uint8_t cls, uint8_t dest, const huff_table *t) { /* :253-269 */
    w_marker(w, 0xC4);
    w_u16(w, (uint16_t)(2 + 1 + 16 + t->n_values));
    w_u8(w, (uint8_t)((cls << 4) | dest));
    w_write(w, t->length, 16);
    w_write(w, t->values, (size_t)t->n_values);
}
static void write_quantization_segment(writer *w, uint8_t dest, const qtable *t) { /* :283-300 */
    w_marker(w, 0xDB);
    w_u16(w, 2 + 1 + 64);
    w_u8(w, dest);
    for (int i = 0; i < 64; i++) w_u8(w, (uint8_t)(t->table[ZIGZAG[i]] >> 3)); /* get(): quantization.rs:285-288 */
}
static void write_frame_header_seg(encoder *e) { /* writer.rs:390-422 */
    writer *w = &e->w;
    w_marker(w, e->p->progressive_scans ? 0xC2 : 0xC0);
    w_u16(w, (uint16_t)(2 + 1 + 2 + 2 + 1 + e->ncomp * 3));
    w_u8(w, 8);
    w_u16(w, e->p->height);
    w_u16(w, e->p->width);
    w_u8(w, (uint8_t)e->ncomp);
    for (int i = 0; i < e->ncomp; i++) {
        w_u8(w, e->comps[i].id);
        w_u8(w, (uint8_t)((e->comps[i].h << 4) | e->comps[i].v));
        w_u8(w, e->comps[i].quantization_table);
    }
}
static void write_scan_header(encoder *e, const component *const *comps, int n, int has_spectral, uint8_t ss, uint8_t se) {
    writer *w = &e->w; /* writer.rs:424-452 */
    w_marker(w, 0xDA);
    w_u16(w, (uint16_t)(2 + 1 + n * 2 + 3));
    w_u8(w, (uint8_t)n);
    for (int i = 0; i < n; i++) {
        w_u8(w, comps[i]->id);
        w_u8(w, (uint8_t)((comps[i]->dc_huffman_table << 4) | comps[i]->ac_huffman_table));
    }
    if (!has_spectral) { ss = 0; se = 63; }
    w_u8(w, ss);
    w_u8(w, se);
    w_u8(w, 0);
}
/* Encoder::write_frame_header, src/encoder.rs:633-667 */
static void write_frame_header(encoder *e) {
    write_frame_header_seg(e);
    write_quantization_segment(&e->w, 0, &e->q[0]);
    write_quantization_segment(&e->w, 1, &e->q[1]);
    write_huffman_segment(&e->w, 0, 0, &e->huff[0][0]);
    write_huffman_segment(&e->w, 1, 0, &e->huff[0][1]);
    if (e->ncomp >= 3) {
        write_huffman_segment(&e->w, 0, 1, &e->huff[1][0]);
        write_huffman_segment(&e->w, 1, 1, &e->huff[1][1]);
    }
    if (e->p->restart_interval) { /* write_dri, writer.rs:302-306 */
        w_marker(&e->w, 0xDD);
        w_u16(&e->w, 4);
        w_u16(&e->w, e->p->restart_interval);
    }
}

/* restart bookkeeping shared by all scan loops (src/encoder.rs:723-725, 748-757, 793-800) */
typedef struct { uint16_t interval, to_go; uint32_t restarts; } restart_state;
static void rs_init(restart_state *r, uint16_t interval) { r->interval = interval; r->to_go = interval; r->restarts = 0; }
static int rs_due(const restart_state *r) { return r->interval > 0 && r->to_go == 0; }
static void rs_step(restart_state *r) {
    if (r->interval > 0) {
        if (r->to_go == 0) { r->to_go = r->interval; r->restarts += 1; r->restarts &= 7; }
        r->to_go -= 1;
    }
}

/* encode_image_interleaved, src/encoder.rs:699-807 */
static int encode_image_interleaved(encoder *e) {
    const orc_params *p = e->p;
    write_frame_header(e);
    const component *all[4];
    for (int i = 0; i < e->ncomp; i++) all[i] = &e->comps[i];
    write_scan_header(e, all, e->ncomp, 0, 0, 0);

    size_t max_h = (size_t)e->max_h, max_v = (size_t)e->max_v;
    size_t width = p->width, height = p->height;
    size_t num_cols = (width + 8 * max_h - 1) / (8 * max_h);
    size_t num_rows = (height + 8 * max_v - 1) / (8 * max_v);
    size_t buffer_width = num_cols * 8 * max_h;
    size_t buffer_size = buffer_width * 8 * max_v;

    uint8_t *row[4] = {0, 0, 0, 0};
    for (int c = 0; c < e->ncomp; c++) {
        row[c] = (uint8_t *)malloc(buffer_size);
        if (!row[c]) { for (int k = 0; k < c; k++) free(row[k]); return ORC_NOMEM; }
    }
    int16_t prev_dc[4] = {0, 0, 0, 0};
    restart_state rs;
    rs_init(&rs, p->restart_interval);

    for (size_t block_y = 0; block_y < num_rows; block_y++) {
        for (size_t yy = 0; yy < 8 * max_v; yy++) { /* :732-745 */
            size_t y = yy + block_y * 8 * max_v;
            if (y > height - 1) y = height - 1;
            uint8_t *dst[4];
            for (int c = 0; c < 4; c++) dst[c] = row[c] ? row[c] + yy * buffer_width : 0;
            fill_row(p, e->data, (uint16_t)y, dst);
            for (int c = 0; c < e->ncomp; c++)
                for (size_t x = width; x < buffer_width; x++) dst[c][x] = dst[c][x - 1];
        }
        for (size_t block_x = 0; block_x < num_cols; block_x++) {
            if (rs_due(&rs)) { /* :748-757 */
                finalize_bit_buffer(&e->w);
                w_marker(&e->w, (uint8_t)(0xD0 + (rs.restarts % 8)));
                prev_dc[0] = prev_dc[1] = prev_dc[2] = prev_dc[3] = 0;
            }
            for (int i = 0; i < e->ncomp; i++) {
                const component *c = &e->comps[i];
                for (size_t v_offset = 0; v_offset < c->v; v_offset++)
                    for (size_t h_offset = 0; h_offset < c->h; h_offset++) {
                        int16_t block[64], q_block[64];
                        get_block(row[i], block_x * 8 * max_h + h_offset * 8, v_offset * 8, max_h / c->h,
                                  max_v / c->v, buffer_width, block);
                        orc_fdct(block);
                        quantize_block(block, q_block, &e->q[c->quantization_table]);
                        write_block(&e->w, q_block, prev_dc[i], &e->huff[c->dc_huffman_table][0],
                                    &e->huff[c->ac_huffman_table][1]);
                        prev_dc[i] = q_block[0];
                    }
            }
            rs_step(&rs);
        }
    }
    finalize_bit_buffer(&e->w);
    for (int c = 0; c < e->ncomp; c++) free(row[c]);
    return ORC_OK;
}

/* encode_blocks, src/encoder.rs:977-1056. blocks[c]: cols_c*rows_c blocks in raster order. */
typedef struct { int16_t *data; size_t n; } block_vec;

static int encode_blocks(encoder *e, block_vec blocks[4]) {
    const orc_params *p = e->p;
    size_t width = p->width, height = p->height;
    size_t max_h = (size_t)e->max_h, max_v = (size_t)e->max_v;
    size_t num_cols = (width + 8 * max_h - 1) / (8 * max_h) * max_h;
    size_t num_rows = (height + 8 * max_v - 1) / (8 * max_v) * max_v;
    size_t buffer_width = num_cols * 8;
    size_t buffer_size = num_cols * num_rows * 64;
    uint8_t *row[4] = {0, 0, 0, 0};
    for (int c = 0; c < 4; c++) { blocks[c].data = 0; blocks[c].n = 0; }
    for (int c = 0; c < e->ncomp; c++) {
        row[c] = (uint8_t *)malloc(buffer_size);
        if (!row[c]) { for (int k = 0; k < c; k++) free(row[k]); return ORC_NOMEM; }
    }
    for (size_t yy = 0; yy < num_rows * 8; yy++) { /* :998-1010 */
        size_t y = yy > height - 1 ? height - 1 : yy;
        uint8_t *dst[4];
        for (int c = 0; c < 4; c++) dst[c] = row[c] ? row[c] + yy * buffer_width : 0;
        fill_row(p, e->data, (uint16_t)y, dst);
        for (int c = 0; c < e->ncomp; c++)
            for (size_t x = width; x < buffer_width; x++) dst[c][x] = dst[c][x - 1];
    }
    num_cols = (width + 7) / 8;
    num_rows = (height + 7) / 8;
    for (int i = 0; i < e->ncomp; i++) { /* :1020-1053 */
        const component *c = &e->comps[i];
        size_t h_scale = max_h / c->h, v_scale = max_v / c->v;
        size_t cols = (num_cols + h_scale - 1) / h_scale, rows = (num_rows + v_scale - 1) / v_scale;
        blocks[i].n = cols * rows;
        blocks[i].data = (int16_t *)malloc(blocks[i].n * 64 * sizeof(int16_t));
        if (!blocks[i].data) return ORC_NOMEM;
        int16_t *out = blocks[i].data;
        for (size_t block_y = 0; block_y < rows; block_y++)
            for (size_t block_x = 0; block_x < cols; block_x++) {
                int16_t block[64];
                get_block(row[i], block_x * 8 * h_scale, block_y * 8 * v_scale, h_scale, v_scale, buffer_width, block);
                orc_fdct(block);
                quantize_block(block, out, &e->q[c->quantization_table]);
                out += 64;
            }
    }
    for (int c = 0; c < e->ncomp; c++) free(row[c]);
    return ORC_OK;
}

/* optimize_huffman_table, src/encoder.rs:1086-1200 */
static int optimize_huffman_table(encoder *e, const block_vec blocks[4]) {
    int max_tables = e->ncomp < 2 ? e->ncomp : 2;
    for (int table = 0; table < max_tables; table++) {
        uint32_t dc_freq[257], ac_freq[257];
        memset(dc_freq, 0, sizeof(dc_freq));
        memset(ac_freq, 0, sizeof(ac_freq));
        dc_freq[256] = 1;
        ac_freq[256] = 1;
        for (int i = 0; i < e->ncomp; i++) {
            const component *c = &e->comps[i];
            if (c->dc_huffman_table == table) {
                int16_t prev_dc = 0;
                for (size_t b = 0; b < blocks[i].n; b++) {
                    int16_t value = blocks[i].data[b * 64];
                    int16_t diff = (int16_t)(value - prev_dc);
                    dc_freq[orc_get_num_bits(diff)] += 1;
                    prev_dc = value;
                }
            }
            if (c->ac_huffman_table == table) {
                int scans = e->p->progressive_scans ? e->p->progressive_scans - 1 : 1;
                int values_per_scan = e->p->progressive_scans ? 64 / scans : 0;
                for (int scan = 0; scan < scans; scan++) {
                    int start, end;
                    if (e->p->progressive_scans) { /* :1122-1134 */
                        start = scan * values_per_scan;
                        if (start < 1) start = 1;
                        end = scan == scans - 1 ? 64 : (scan + 1) * values_per_scan;
                    } else { /* :1163-1187 */
                        start = 1;
                        end = 64;
                    }
                    for (size_t b = 0; b < blocks[i].n; b++) {
                        const int16_t *block = blocks[i].data + b * 64;
                        uint32_t zero_run = 0;
                        for (int k = start; k < end; k++) {
                            int16_t value = block[k];
                            if (value == 0) {
                                zero_run += 1;
                            } else {
                                while (zero_run > 15) { ac_freq[0xF0] += 1; zero_run -= 16; }
                                uint32_t symbol = (zero_run << 4) | orc_get_num_bits(value);
                                ac_freq[symbol] += 1;
                                zero_run = 0;
                            }
                        }
                        if (zero_run > 0) ac_freq[0] += 1;
                    }
                }
            }
        }
        huff_table *dc = &e->huff[table][0], *ac = &e->huff[table][1];
        dc->n_values = orc_huffman_optimized(dc_freq, dc->length, dc->values);
        ac->n_values = orc_huffman_optimized(ac_freq, ac->length, ac->values);
        if (dc->n_values < 0 || ac->n_values < 0) return ORC_BAD_PARAMS;
        huff_build_lookup(dc);
        huff_build_lookup(ac);
    }
    return ORC_OK;
}

static void free_blocks(block_vec blocks[4]) {
    for (int c = 0; c < 4; c++) free(blocks[c].data);
}

/* encode_image_sequential, src/encoder.rs:810-864 */
static int encode_image_sequential(encoder *e) {
    block_vec blocks[4];
    int rc = encode_blocks(e, blocks);
    if (rc == ORC_OK && e->p->optimize_huffman) rc = optimize_huffman_table(e, blocks);
    if (rc != ORC_OK) { free_blocks(blocks); return rc; }
    write_frame_header(e);
    for (int i = 0; i < e->ncomp; i++) {
        const component *c = &e->comps[i];
        restart_state rs;
        rs_init(&rs, e->p->restart_interval);
        write_scan_header(e, &c, 1, 0, 0, 0);
        int16_t prev_dc = 0;
        for (size_t b = 0; b < blocks[i].n; b++) {
            const int16_t *block = blocks[i].data + b * 64;
            if (rs_due(&rs)) {
                finalize_bit_buffer(&e->w);
                w_marker(&e->w, (uint8_t)(0xD0 + (rs.restarts % 8)));
                prev_dc = 0;
            }
            write_block(&e->w, block, prev_dc, &e->huff[c->dc_huffman_table][0], &e->huff[c->ac_huffman_table][1]);
            prev_dc = block[0];
            rs_step(&rs);
        }
        finalize_bit_buffer(&e->w);
    }
    free_blocks(blocks);
    return ORC_OK;
}

/* encode_image_progressive, src/encoder.rs:869-975 (spectral selection only) */
static int encode_image_progressive(encoder *e) {
    block_vec blocks[4];
    int rc = encode_blocks(e, blocks);
    if (rc == ORC_OK && e->p->optimize_huffman) rc = optimize_huffman_table(e, blocks);
    if (rc != ORC_OK) { free_blocks(blocks); return rc; }
    write_frame_header(e);
    for (int i = 0; i < e->ncomp; i++) { /* phase 1: DC scans :885-922 */
        const component *c = &e->comps[i];
        write_scan_header(e, &c, 1, 1, 0, 0);
        restart_state rs;
        rs_init(&rs, e->p->restart_interval);
        int16_t prev_dc = 0;
        for (size_t b = 0; b < blocks[i].n; b++) {
            const int16_t *block = blocks[i].data + b * 64;
            if (rs_due(&rs)) {
                finalize_bit_buffer(&e->w);
                w_marker(&e->w, (uint8_t)(0xD0 + (rs.restarts % 8)));
                prev_dc = 0;
            }
            write_dc(&e->w, block[0], prev_dc, &e->huff[c->dc_huffman_table][0]);
            prev_dc = block[0];
            rs_step(&rs);
        }
        finalize_bit_buffer(&e->w);
    }
    int scans = e->p->progressive_scans - 1; /* phase 2: AC scans :925-972 */
    int values_per_scan = 64 / scans;
    for (int scan = 0; scan < scans; scan++) {
        int start = scan * values_per_scan;
        if (start < 1) start = 1;
        int end = scan == scans - 1 ? 64 : (scan + 1) * values_per_scan;
        for (int i = 0; i < e->ncomp; i++) {
            const component *c = &e->comps[i];
            restart_state rs;
            rs_init(&rs, e->p->restart_interval);
            write_scan_header(e, &c, 1, 1, (uint8_t)start, (uint8_t)(end - 1));
            for (size_t b = 0; b < blocks[i].n; b++) {
                const int16_t *block = blocks[i].data + b * 64;
                if (rs_due(&rs)) {
                    finalize_bit_buffer(&e->w);
                    w_marker(&e->w, (uint8_t)(0xD0 + (rs.restarts % 8)));
                }
                write_ac_block(&e->w, block, start, end, &e->huff[c->ac_huffman_table][1]);
                rs_step(&rs);
            }
            finalize_bit_buffer(&e->w);
        }
    }
    free_blocks(blocks);
    return ORC_OK;
}

static int check_params(const orc_params *p) {
    if (p->color_type > ORC_YCCK) return ORC_BAD_PARAMS;
    uint8_t h = (p->sampling >> 4) & 7, v = p->sampling & 0xf;
    int ok = (h == 1 || h == 2 || h == 4) && (v == 1 || v == 2 || v == 4) && !(h == 4 && v == 4);
    if (!ok) return ORC_BAD_PARAMS;
    if (p->progressive_scans == 1 || p->progressive_scans > 64) return ORC_BAD_PARAMS; /* encoder.rs:329-333 */
    for (uint32_t i = 0; i < p->n_app; i++)
        if (p->apps[i].nr == 0 || p->apps[i].nr > 15 || p->apps[i].len > 65533) return ORC_BAD_PARAMS; /* :374-383 */
    return ORC_OK;
}

static void encoder_init(encoder *e, const orc_params *p, const uint8_t *pixels) {
    memset(e, 0, sizeof(*e));
    e->p = p;
    e->data = pixels;
    e->w.free_bits = 64;
    huff_new(&e->huff[0][0], LUMA_DC_LEN, DC_VALUES, 12); /* Encoder::new, encoder.rs:239-249 */
    huff_new(&e->huff[0][1], LUMA_AC_LEN, LUMA_AC_VALUES, 162);
    huff_new(&e->huff[1][0], CHROMA_DC_LEN, DC_VALUES, 12);
    huff_new(&e->huff[1][1], CHROMA_AC_LEN, CHROMA_AC_VALUES, 162);
    /* encode_image_internal :528-531 */
    orc_quant_table(p->qtable_kind[0], p->qtable_custom[0], p->quality, 1, e->q[0].table, e->q[0].recip, e->q[0].corr);
    orc_quant_table(p->qtable_kind[1], p->qtable_custom[1], p->quality, 0, e->q[1].table, e->q[1].recip, e->q[1].corr);
    init_components(e);
}

/* Encoder::encode + encode_image_internal, src/encoder.rs:440-567 */
int orc_encode(const orc_params *p, const uint8_t *pixels, size_t len, uint8_t **out, size_t *out_len) {
    *out = 0;
    *out_len = 0;
    int rc = check_params(p);
    if (rc != ORC_OK) return rc;
    size_t required = (size_t)p->width * p->height * (size_t)bytes_per_pixel(p->color_type);
    if (len < required) return ORC_BAD_IMAGE_DATA;                /* :447-454 */
    if (p->width == 0 || p->height == 0) return ORC_ZERO_DIMENSIONS; /* :521-526 */

    encoder *e = (encoder *)malloc(sizeof(encoder));
    if (!e) return ORC_NOMEM;
    encoder_init(e, p, pixels);

    w_marker(&e->w, 0xD8);      /* SOI :536 */
    write_header(&e->w, p);     /* :538 */
    int nc = num_components(p->color_type);
    if (nc == 4 && p->color_type == ORC_CMYK) { /* :540-550 */
        write_segment(&e->w, 0xEE, (const uint8_t *)"Adobe\0\0\0\0\0\0\0", 12);
    } else if (nc == 4) {
        write_segment(&e->w, 0xEE, (const uint8_t *)"Adobe\0\0\0\0\0\0\x02", 12);
    }
    for (uint32_t i = 0; i < p->n_app; i++) /* :552-554 */
        write_segment(&e->w, (uint8_t)(0xE0 + p->apps[i].nr), p->apps[i].data, p->apps[i].len);

    if (p->progressive_scans) rc = encode_image_progressive(e);            /* :556-562 */
    else if (p->optimize_huffman || !supports_interleaved(p->sampling)) rc = encode_image_sequential(e);
    else rc = encode_image_interleaved(e);

    w_marker(&e->w, 0xD9); /* EOI :564 */
    if (rc == ORC_OK && e->w.oom) rc = ORC_NOMEM;
    if (rc != ORC_OK) {
        free(e->w.buf);
    } else {
        *out = e->w.buf;
        *out_len = e->w.len;
    }
    free(e);
    return rc;
}

void orc_free(void *p) { free(p); }

/* Coefficients over the MCU-padded block grid (see header). Same samples / arithmetic as
 * get_block -> fdct -> quantize_block in both walks. */
int orc_coefficients(const orc_params *p, const uint8_t *pixels, size_t len, int16_t *blocks[4], uint32_t n_blocks[4]) {
    for (int c = 0; c < 4; c++) { blocks[c] = 0; n_blocks[c] = 0; }
    int rc = check_params(p);
    if (rc != ORC_OK) return rc;
    size_t required = (size_t)p->width * p->height * (size_t)bytes_per_pixel(p->color_type);
    if (len < required) return ORC_BAD_IMAGE_DATA;
    if (p->width == 0 || p->height == 0) return ORC_ZERO_DIMENSIONS;
    encoder *e = (encoder *)malloc(sizeof(encoder));
    if (!e) return ORC_NOMEM;
    encoder_init(e, p, pixels);
    size_t width = p->width, height = p->height;
    size_t max_h = (size_t)e->max_h, max_v = (size_t)e->max_v;
    size_t mcu_cols = (width + 8 * max_h - 1) / (8 * max_h), mcu_rows = (height + 8 * max_v - 1) / (8 * max_v);
    size_t buffer_width = mcu_cols * 8 * max_h, buffer_height = mcu_rows * 8 * max_v;
    uint8_t *row[4] = {0, 0, 0, 0};
    for (int c = 0; c < e->ncomp; c++) row[c] = (uint8_t *)malloc(buffer_width * buffer_height);
    for (size_t yy = 0; yy < buffer_height; yy++) {
        size_t y = yy > height - 1 ? height - 1 : yy;
        uint8_t *dst[4];
        for (int c = 0; c < 4; c++) dst[c] = row[c] ? row[c] + yy * buffer_width : 0;
        fill_row(p, pixels, (uint16_t)y, dst);
        for (int c = 0; c < e->ncomp; c++)
            for (size_t x = width; x < buffer_width; x++) dst[c][x] = dst[c][x - 1];
    }
    for (int i = 0; i < e->ncomp; i++) {
        const component *c = &e->comps[i];
        size_t h_scale = max_h / c->h, v_scale = max_v / c->v;
        size_t cols = mcu_cols * c->h, rows = mcu_rows * c->v;
        n_blocks[i] = (uint32_t)(cols * rows);
        blocks[i] = (int16_t *)malloc((size_t)n_blocks[i] * 64 * sizeof(int16_t));
        int16_t *o = blocks[i];
        for (size_t by = 0; by < rows; by++)
            for (size_t bx = 0; bx < cols; bx++) {
                int16_t block[64];
                get_block(row[i], bx * 8 * h_scale, by * 8 * v_scale, h_scale, v_scale, buffer_width, block);
                orc_fdct(block);
                quantize_block(block, o, &e->q[c->quantization_table]);
                o += 64;
            }
    }
    for (int c = 0; c < e->ncomp; c++) free(row[c]);
    free(e);
    return ORC_OK;
}
