// Device-wide exclusive prefix sum (u32 in, u64 out): the "device-wide prefix sum of per-block bit
// lengths" that turns the reference's serial bit writer (src/writer.rs:186-202) into independent
// placements. Three launches: per-tile reduce, one-CTA scan of the tile sums, per-tile downsweep.
#include "kernels.h"

namespace jpgb {
namespace {

constexpr int kThreads = 256;
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;

__device__ __forceinline__ unsigned long long warp_inclusive(unsigned long long v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// exclusive scan of one value per thread across the CTA; returns the CTA total through `total`
template <int THREADS>
__device__ __forceinline__ unsigned long long block_exclusive(unsigned long long v, unsigned long long *smem /*THREADS/32+1*/,
                                                              unsigned long long &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long inc = warp_inclusive(v, lane);
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = lane < THREADS / 32 ? smem[lane] : 0ull;
        const unsigned long long winc = warp_inclusive(w, lane);
        if (lane < THREADS / 32) smem[lane] = winc - w;
        if (lane == 31) smem[THREADS / 32] = winc;
    }
    __syncthreads();
    const unsigned long long r = smem[warp] + inc - v;
    total = smem[THREADS / 32];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kThreads) tile_reduce_kernel(const uint32_t *__restrict__ in, unsigned long long n,
                                                               unsigned long long *__restrict__ tile_sums) {
    __shared__ unsigned long long sm[kThreads / 32 + 1];
    const unsigned long long base = (unsigned long long)blockIdx.x * kTile;
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const unsigned long long i = base + (unsigned long long)k * kThreads + threadIdx.x;
        if (i < n) s += in[i];
    }
    unsigned long long total;
    block_exclusive<kThreads>(s, sm, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_sums_kernel(unsigned long long *tile_sums, unsigned long long n_tiles,
                                                         unsigned long long *total_out) {
    __shared__ unsigned long long sm[1024 / 32 + 1];
    unsigned long long carry = 0;
    for (unsigned long long base = 0; base < n_tiles; base += 1024) {
        const unsigned long long i = base + threadIdx.x;
        const unsigned long long v = i < n_tiles ? tile_sums[i] : 0ull;
        unsigned long long total;
        const unsigned long long ex = block_exclusive<1024>(v, sm, total);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(kThreads) tile_downsweep_kernel(const uint32_t *__restrict__ in, unsigned long long n,
                                                                  const unsigned long long *__restrict__ tile_offs,
                                                                  unsigned long long *__restrict__ out) {
    __shared__ unsigned long long sm[kThreads / 32 + 1];
    // each thread owns kItems consecutive elements
    const unsigned long long base = (unsigned long long)blockIdx.x * kTile + (unsigned long long)threadIdx.x * kItems;
    uint32_t v[kItems];
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        v[k] = base + k < n ? in[base + k] : 0u;
        s += v[k];
    }
    unsigned long long total;
    unsigned long long run = block_exclusive<kThreads>(s, sm, total) + tile_offs[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

} // namespace

size_t scan_tmp_bytes(uint64_t n) { return ((n + kTile - 1) / kTile + 1) * sizeof(unsigned long long); }

cudaError_t launch_exclusive_scan(const uint32_t *in, unsigned long long *out, uint64_t n, void *tmp, cudaStream_t stream,
                                  uint32_t *launches) {
    const uint64_t tiles = (n + kTile - 1) / kTile;
    auto *sums = static_cast<unsigned long long *>(tmp);
    if (tiles > 0) tile_reduce_kernel<<<(unsigned)tiles, kThreads, 0, stream>>>(in, n, sums);
    scan_sums_kernel<<<1, 1024, 0, stream>>>(sums, tiles, out + n);
    if (tiles > 0) tile_downsweep_kernel<<<(unsigned)tiles, kThreads, 0, stream>>>(in, n, sums, out);
    if (launches) *launches += tiles > 0 ? 3 : 1;
    return cudaGetLastError();
}

} // namespace jpgb
