// Stage A: packed pixels -> quantized zig-zag coefficients, one fused kernel.
//
// Replaces, bit for bit, the reference's per-block numeric chain
//   ImageBuffer::fill_buffers (colour conversion)       src/image_buffer.rs:9-38, 100-313
//   row padding (replicate last sample / last row)       src/encoder.rs:732-745, 998-1010
//   get_block (strided gather = point decimation, -128)  src/encoder.rs:1222-1242
//   fdct (jfdctint islow, i32, x8 scaled)                src/fdct.rs:107-238
//   Operations::quantize_block (reciprocal, zig-zag)     src/encoder.rs:1265-1271, quantization.rs:291-307
//
// Work decomposition: a CTA owns a tile of `groups` x 32 MCUs of one MCU row. The tile's pixel rows
// are staged into shared memory with 128-bit coalesced loads (edge pixels replicated while staging,
// so every later read is in-bounds). A *warp task* is 32 blocks with the same (component, v, h)
// position in 32 consecutive MCUs: each lane owns one whole 8x8 block in registers, so both DCT
// passes, the transpose between them and the zig-zag permutation are register renaming -- no
// shuffles, no shared-memory round trip, no divergence. Each lane then writes its 128-byte block.
//
// All arithmetic is 32-bit integer; tensor cores are not used (the DCT must be bit-exact).
#include "kernels.h"

namespace jpgb {

namespace {

__device__ __forceinline__ int clamp_u8_formula_y(int r, int g, int b) {
    return (19595 * r + 38470 * g + 7471 * b + 0x7FFF) >> 16; // image_buffer.rs:22,26
}
__device__ __forceinline__ int formula_cb(int r, int g, int b) {
    return (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 0x7FFF) >> 16; // :23,27
}
__device__ __forceinline__ int formula_cr(int r, int g, int b) {
    return (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 0x7FFF) >> 16; // :24,28
}

// One sample of component `comp` from the pixel at `px` (shared memory), unshifted 0..255.
template <int CT>
__device__ __forceinline__ int sample(const uint8_t *px, int comp) {
    if (CT == JPGB_LUMA) return px[0];
    if (CT == JPGB_YCBCR || CT == JPGB_YCCK) return px[comp];
    if (CT == JPGB_CMYK) return 255 - px[comp];
    int r, g, b;
    if (CT == JPGB_BGR || CT == JPGB_BGRA) {
        r = px[2]; g = px[1]; b = px[0];
    } else {
        r = px[0]; g = px[1]; b = px[2];
    }
    if (CT == JPGB_CMYK_AS_YCCK && comp == 3) return 255 - px[3];
    if (comp == 0) return clamp_u8_formula_y(r, g, b);
    if (comp == 1) return formula_cb(r, g, b);
    return formula_cr(r, g, b);
}

// 1-D 8-point LL&M forward DCT (fdct.rs:116-171 for PASS 1, :178-237 for PASS 2).
// PASS 1 takes *unshifted* samples 0..255: the -128 level shift (encoder.rs:1237) only moves the
// DC term of the row by -128*8 << PASS1_BITS, every other output is a function of differences.
template <int PASS>
__device__ __forceinline__ void dct8(int &d0, int &d1, int &d2, int &d3, int &d4, int &d5, int &d6, int &d7) {
    constexpr int N = PASS == 1 ? 11 : 15; // CONST_BITS -/+ PASS1_BITS
    constexpr int RND = 1 << (N - 1);
    const int tmp0 = d0 + d7, tmp7 = d0 - d7;
    const int tmp1 = d1 + d6, tmp6 = d1 - d6;
    const int tmp2 = d2 + d5, tmp5 = d2 - d5;
    const int tmp3 = d3 + d4, tmp4 = d3 - d4;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3;
    const int tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    if (PASS == 1) {
        d0 = (tmp10 + tmp11 - 1024) << 2;
        d4 = (tmp10 - tmp11) << 2;
    } else {
        d0 = (tmp10 + tmp11 + 2) >> 2;
        d4 = (tmp10 - tmp11 + 2) >> 2;
    }
    const int z1e = (tmp12 + tmp13) * 4433 + RND;      // FIX_0_541196100
    d2 = (z1e + tmp13 * 6270) >> N;                    // FIX_0_765366865
    d6 = (z1e - tmp12 * 15137) >> N;                   // FIX_1_847759065
    const int z1 = tmp4 + tmp7, z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
    const int z5 = (z3 + z4) * 9633 + RND;             // FIX_1_175875602
    const int z3m = z5 - z3 * 16069;                   // FIX_1_961570560
    const int z4m = z5 - z4 * 3196;                    // FIX_0_390180644
    const int z1m = z1 * -7373;                        // FIX_0_899976223
    const int z2m = z2 * -20995;                       // FIX_2_562915447
    d7 = (tmp4 * 2446 + z1m + z3m) >> N;               // FIX_0_298631336
    d5 = (tmp5 * 16819 + z2m + z4m) >> N;              // FIX_2_053119869
    d3 = (tmp6 * 25172 + z2m + z3m) >> N;              // FIX_3_072711026
    d1 = (tmp7 * 12299 + z1m + z4m) >> N;              // FIX_1_501321110
}

struct ZZ {
    int v[64];
};
__host__ __device__ constexpr ZZ make_zz() {
    return ZZ{{0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
               41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
               30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63}};
}

// quantize natural-order coefficient n with table T: the upper half of the 32-bit result is the
// quantized value (see QuantConsts in device_types.h)
template <int T>
__device__ __forceinline__ int quant32(const StageAParams &p, int v, int n) {
    const int add = v < 0 ? p.q[T].add_neg[n] : p.q[T].add_pos[n];
    return v * p.q[T].mul[n] + add;
}

template <int T>
__device__ __forceinline__ void quantize_store(const StageAParams &p, const int (&v)[64], int16_t *dst) {
    constexpr ZZ zz = make_zz();
    uint4 *out = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        uint32_t r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i0 = w * 8 + k * 2, i1 = i0 + 1;
            const int a = quant32<T>(p, v[zz.v[i0]], zz.v[i0]);
            const int b = quant32<T>(p, v[zz.v[i1]], zz.v[i1]);
            r[k] = __byte_perm((uint32_t)a, (uint32_t)b, 0x7632); // {hi16(a), hi16(b)}
        }
        out[w] = make_uint4(r[0], r[1], r[2], r[3]);
    }
}

template <int CT>
__global__ void __launch_bounds__(256) stage_a_kernel(const __grid_constant__ StageAParams p) {
    extern __shared__ __align__(16) uint8_t tile[];
    constexpr int BPP = CT == JPGB_LUMA ? 1 : ((CT == JPGB_RGB || CT == JPGB_BGR || CT == JPGB_YCBCR) ? 3 : 4);

    const int tile_x = blockIdx.x, mcu_y = blockIdx.y, img = blockIdx.z;
    const int mcu_x0 = tile_x * 32 * p.groups;
    const int px0 = mcu_x0 * 8 * p.hmax, py0 = mcu_y * 8 * p.vmax;
    const uint8_t *src = p.pixels + (size_t)img * p.image_stride;
    const size_t row_bytes = (size_t)p.width * BPP;

    // ---- stage the tile: 16-byte chunks, edges replicated (Q4) ----
    const int chunks_per_row = p.tile_pitch / 16;
    const int n_chunks = chunks_per_row * p.tile_h_px;
    const int valid_px = min(p.tile_w_px, p.width - px0);  // > 0 by construction
    const int valid_bytes = valid_px * BPP;
    for (int c = threadIdx.x; c < n_chunks; c += blockDim.x) {
        const int ry = c / chunks_per_row, cb = (c - ry * chunks_per_row) * 16;
        const int sy = min(py0 + ry, p.height - 1);
        const uint8_t *row = src + (size_t)sy * row_bytes + (size_t)px0 * BPP;
        uint8_t *dst = tile + ry * p.tile_pitch + cb;
        if (cb + 16 <= valid_bytes && ((reinterpret_cast<uintptr_t>(row + cb) & 15) == 0)) {
            *reinterpret_cast<uint4 *>(dst) = __ldg(reinterpret_cast<const uint4 *>(row + cb));
        } else {
#pragma unroll 4
            for (int b = 0; b < 16; ++b) {
                const int byte = cb + b;
                const int px = byte / BPP, ch = byte - px * BPP;
                dst[b] = row[min(px, valid_px - 1) * BPP + ch];
            }
        }
    }
    __syncthreads();

    // ---- warp tasks ----
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int n_tasks = p.groups * p.tasks_per_group;
    for (int task = warp; task < n_tasks; task += n_warps) {
        const int group = task / p.tasks_per_group, slot = task - group * p.tasks_per_group;
        const int comp = p.task_comp[slot], bv = p.task_v[slot], bh = p.task_h[slot];
        const int mcu_local = group * 32 + lane;
        const int mcu_x = mcu_x0 + mcu_local;
        if (mcu_x >= p.mcu_cols) continue;
        // get_block arguments, encoder.rs:762-769: start = mcu*8*max + offset*8, stride = max/factor
        const int sx = p.hmax / p.comp_h[comp], sy = p.vmax / p.comp_v[comp];
        const int x0 = mcu_local * 8 * p.hmax + bh * 8, y0 = bv * 8;
        const uint8_t *base = tile + y0 * p.tile_pitch + x0 * BPP;
        const int step_x = sx * BPP, step_y = sy * p.tile_pitch;

        int v[64];
#pragma unroll
        for (int y = 0; y < 8; ++y)
#pragma unroll
            for (int x = 0; x < 8; ++x) v[y * 8 + x] = sample<CT>(base + y * step_y + x * step_x, comp);

#pragma unroll
        for (int y = 0; y < 8; ++y)
            dct8<1>(v[y * 8 + 0], v[y * 8 + 1], v[y * 8 + 2], v[y * 8 + 3], v[y * 8 + 4], v[y * 8 + 5], v[y * 8 + 6], v[y * 8 + 7]);
#pragma unroll
        for (int x = 0; x < 8; ++x)
            dct8<2>(v[x], v[8 + x], v[16 + x], v[24 + x], v[32 + x], v[40 + x], v[48 + x], v[56 + x]);

        const size_t blk = (size_t)img * p.blocks_per_image + p.comp_off[comp] +
                           (size_t)(mcu_y * p.comp_v[comp] + bv) * p.comp_pw[comp] + (size_t)mcu_x * p.comp_h[comp] + bh;
        int16_t *dst = p.coef + blk * 64;
        if (p.comp_qt[comp] == 0) quantize_store<0>(p, v, dst);
        else quantize_store<1>(p, v, dst);
    }
}

} // namespace

cudaError_t launch_stage_a(const StageAParams &p, uint32_t n_images, cudaStream_t stream) {
    const size_t smem = (size_t)p.tile_pitch * p.tile_h_px;
    const int n_tasks = p.groups * p.tasks_per_group;
    const int warps = n_tasks < 8 ? n_tasks : 8;
    dim3 grid(p.tiles_per_row, p.mcu_rows, n_images), block(warps * 32);
#define JPGB_LAUNCH_A(CT)                                                                                        \
    case CT: {                                                                                                   \
        if (smem > 48 * 1024) {                                                                                  \
            cudaError_t e = cudaFuncSetAttribute(stage_a_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return e;                                                                      \
        }                                                                                                        \
        stage_a_kernel<CT><<<grid, block, smem, stream>>>(p);                                                    \
        break;                                                                                                   \
    }
    switch (p.color_type) {
        JPGB_LAUNCH_A(JPGB_LUMA)
        JPGB_LAUNCH_A(JPGB_RGB)
        JPGB_LAUNCH_A(JPGB_RGBA)
        JPGB_LAUNCH_A(JPGB_BGR)
        JPGB_LAUNCH_A(JPGB_BGRA)
        JPGB_LAUNCH_A(JPGB_YCBCR)
        JPGB_LAUNCH_A(JPGB_CMYK)
        JPGB_LAUNCH_A(JPGB_CMYK_AS_YCCK)
        JPGB_LAUNCH_A(JPGB_YCCK)
    default: return cudaErrorInvalidValue;
    }
#undef JPGB_LAUNCH_A
    return cudaGetLastError();
}

} // namespace jpgb
