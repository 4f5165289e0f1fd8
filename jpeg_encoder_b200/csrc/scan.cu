// Device-wide exclusive prefix sum (u32 in, u64 out): the "device-wide prefix sum of per-block bit
// lengths" that turns the reference's serial bit writer (src/writer.rs:186-202) into independent
// placements.
//
// Large inputs: one pass with decoupled look-back. A CTA takes the next tile from a ticket counter
// (so every predecessor of a running tile has started), scans its 2048 elements in registers,
// publishes its aggregate, then warp 0 looks back over the predecessors' descriptors (32 at a time)
// until it meets an inclusive prefix, publishes its own inclusive prefix and the CTA writes
// prefix + local scan. The input is read once and the output written once.
// Small inputs (<= 4096 elements): a single CTA, one launch.
#include "kernels.h"

namespace jpgb {
namespace {

constexpr int kThreads = 256;
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;
constexpr uint64_t kSmallN = 4096;

// descriptor: bits 63..62 status, bits 61..0 value
constexpr unsigned long long kStatusAggregate = 1ull << 62, kStatusPrefix = 2ull << 62, kValueMask = (1ull << 62) - 1;

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

__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// state[0] = ticket counter, state[2 + t] = descriptor of tile t (all zero on entry); *err_flag is raised if a
// predecessor never publishes (cannot happen with ticketed tiles; the bounded spin only guards the device)
__global__ void __launch_bounds__(kThreads) lookback_scan_kernel(const uint32_t *__restrict__ in, unsigned long long n,
                                                                 unsigned long long *__restrict__ out, unsigned long long *state,
                                                                 unsigned long long *err_flag) {
    __shared__ unsigned long long sm[kThreads / 32 + 1];
    __shared__ unsigned long long s_tile, s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(state, 1ull);
    __syncthreads();
    const unsigned long long tile = s_tile;
    unsigned long long *desc = state + 2;

    // each thread owns kItems consecutive elements
    const unsigned long long base = tile * kTile + (unsigned long long)threadIdx.x * kItems;
    uint32_t v[kItems];
    if (base + kItems <= n) {
        const uint4 a = *reinterpret_cast<const uint4 *>(in + base), b = *reinterpret_cast<const uint4 *>(in + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < kItems; ++k) v[k] = base + k < n ? in[base + k] : 0u;
    }
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) s += v[k];
    unsigned long long total;
    const unsigned long long local = block_exclusive<kThreads>(s, sm, total);

    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        if (lane == 0) st_release(desc + tile, (tile == 0 ? kStatusPrefix : kStatusAggregate) | total);
        unsigned long long prefix = 0;
        if (tile > 0) {
            long long look = (long long)tile - 1; // lane j inspects tile look - j
            unsigned spins = 0;
            for (;;) {
                const long long t = look - lane;
                unsigned long long d = t >= 0 ? ld_acquire(desc + t) : kStatusPrefix; // before tile 0: an empty prefix
                // every lane needs a published descriptor up to the first inclusive prefix
                const unsigned ready = __ballot_sync(0xffffffffu, (d >> 62) != 0);
                const unsigned is_prefix = __ballot_sync(0xffffffffu, (d >> 62) == 2);
                const int first_prefix = is_prefix ? __ffs((int)is_prefix) - 1 : 32; // nearest tile carrying an inclusive prefix
                const unsigned need = first_prefix >= 31 ? 0xffffffffu : ((2u << first_prefix) - 1u);
                if ((ready & need) != need) { // a predecessor has not published yet
                    if (++spins > (1u << 26)) { // safety net: never hang the device
                        if (lane == 0 && err_flag) atomicExch(err_flag, 1ull);
                        break;
                    }
                    __nanosleep(20);
                    continue;
                }
                unsigned long long contrib = (lane <= first_prefix) ? (d & kValueMask) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                prefix += contrib;
                if (first_prefix < 32) break;
                look -= 32;
            }
            if (lane == 0) st_release(desc + tile, kStatusPrefix | ((prefix + total) & kValueMask));
        }
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    unsigned long long run = s_prefix + local;
    if (base + kItems <= n) {
        unsigned long long o[kItems];
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            o[k] = run;
            run += v[k];
        }
        ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(out + base);
#pragma unroll
        for (int k = 0; k < kItems / 2; ++k) dst[k] = make_ulonglong2(o[2 * k], o[2 * k + 1]);
    } else {
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            if (base + k < n) out[base + k] = run;
            run += v[k];
        }
    }
    if (tile == (n + kTile - 1) / kTile - 1 && threadIdx.x == 0) out[n] = s_prefix + total; // grand total
}

__global__ void __launch_bounds__(1024) small_scan_kernel(const uint32_t *__restrict__ in, unsigned long long n,
                                                          unsigned long long *__restrict__ out) {
    __shared__ unsigned long long sm[1024 / 32 + 1];
    unsigned long long carry = 0;
    for (unsigned long long base = 0; base < n; base += 1024) {
        const unsigned long long i = base + threadIdx.x;
        const unsigned long long v = i < n ? in[i] : 0ull;
        unsigned long long total;
        const unsigned long long ex = block_exclusive<1024>(v, sm, total);
        if (i < n) out[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) out[n] = carry;
}

} // namespace

// scratch: ticket + error flag + one descriptor per tile
size_t scan_tmp_bytes(uint64_t n) { return ((n + kTile - 1) / kTile + 4) * sizeof(unsigned long long); }

cudaError_t launch_exclusive_scan(const uint32_t *in, unsigned long long *out, uint64_t n, void *tmp, cudaStream_t stream,
                                  uint32_t *launches, unsigned long long *err_flag) {
    if (n <= kSmallN) {
        small_scan_kernel<<<1, 1024, 0, stream>>>(in, n, out);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    const uint64_t tiles = (n + kTile - 1) / kTile;
    cudaError_t e = cudaMemsetAsync(tmp, 0, (tiles + 2) * sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    lookback_scan_kernel<<<(unsigned)tiles, kThreads, 0, stream>>>(in, n, out, static_cast<unsigned long long *>(tmp), err_flag);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

} // namespace jpgb
