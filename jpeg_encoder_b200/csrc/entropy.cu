// Stage B: quantized coefficients -> JFIF bytes, entirely on the device.
//
// Replaces the reference's serial entropy coder and bit writer
//   write_dc / write_ac_block / write_block / get_code      src/writer.rs:331-388, 455-470
//   write_bits / flush / finalize_bit_buffer (0xFF stuffing) src/writer.rs:138-202
//   the MCU / block walks with restart bookkeeping           src/encoder.rs:727-804, 823-861, 885-972
//   optimize_huffman_table's symbol histogram                src/encoder.rs:1086-1200
//
// Vocabulary. A *visit* is one block coded in one scan (a block is visited once per scan that
// touches it); visits are numbered in the reference's emission order. A *segment* is the run of
// visits between two restart points of one scan (one segment per scan when restarts are off);
// its bits start byte-aligned and end padded with 1-bits. The *unstuffed stream* is the complete
// file before 0xFF stuffing: per image the header, then per segment its lead (RSTn marker or the
// next scan's SOS) and its data bytes, then EOI. `raw_mask` flags header/marker bytes so that the
// stuffing pass leaves their 0xFF alone.
//
//   encode_visits_kernel visit -> its code bits (in a slot) and their count (DC differencing, run/size symbols)
//   [exclusive scan]     bit position of every visit
//   segment_len_kernel   segment -> lead + ceil(bits/8) + tail bytes
//   [exclusive scan]     byte position of every segment in the unstuffed stream
//   segment_lead_kernel  writes headers / RSTn / SOS / EOI, sets raw_mask
//   place_bits_kernel    slot bits shifted into the unstuffed stream, pad bits at segment end
//   count_ff_kernel      chunk -> number of data 0xFF bytes
//   [exclusive scan]
//   stuff_scatter_kernel copies every byte to its final place, inserting 0x00 after data 0xFF
#include "kernels.h"

namespace jpgb {
namespace {

struct VisitInfo {
    const int16_t *blk;   // this block's 64 zig-zag coefficients
    const int16_t *pred;  // block holding the DC predictor, or nullptr for "predictor is 0"
    int comp, ss, se, tbl;
    int pred_back;        // the predecessor block is the block of visit (this - pred_back)
    unsigned long long first_visit_of_seg; // within the image
    unsigned seg_local;                    // segment index within the image
    unsigned seg_in_scan;
    int scan;
    bool last_of_seg;
};

__device__ __forceinline__ int find_scan_by_visit(const DevPlan &P, unsigned long long v) {
    int lo = 0, hi = P.n_scans - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (P.scans[mid].visit_base <= v) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}
__device__ __forceinline__ int find_scan_by_seg(const DevPlan &P, unsigned s) {
    int lo = 0, hi = P.n_scans - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (P.scans[mid].seg_base <= s) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

// g = img * visits_per_image + v. Nearly every call has both numbers below 2^32 (FastDiv); the 64-bit
// divide is ~100 instructions.
__device__ __forceinline__ void split_visit(const DevPlan &P, unsigned long long g, unsigned long long &img, unsigned long long &v) {
    if ((g >> 32) == 0 && P.div_vpi.d != 0) {
        unsigned q, r;
        divmod((unsigned)g, P.div_vpi, q, r);
        img = q;
        v = r;
    } else {
        img = g / P.visits_per_image;
        v = g - img * P.visits_per_image;
    }
}
// unit and slot of a visit inside its scan; a scan has < 2^32 visits (<= 2^26 blocks x 18 per unit)
__device__ __forceinline__ void split_unit(const DevScan &S, unsigned long long v, unsigned &unit, unsigned &slot) {
    const unsigned rel = (unsigned)(v - S.visit_base);
    if (S.bpu == 1) {
        unit = rel;
        slot = 0;
    } else {
        divmod(rel, S.div_bpu, unit, slot);
    }
}

// Where does visit `v` (numbered within one image) live, and what precedes it?
// Interleaved order: encoder.rs:747-791 (MCU raster; component, v, h inside the MCU).
// Single-component order: encoder.rs:832 / 894 / 946 over encode_blocks' raster grid (:1030-1031).
// Predictor reset at restart points: :753-756, :838, :900.
__device__ __forceinline__ VisitInfo locate_visit(const DevPlan &P, const int16_t *coef_img, unsigned long long v) {
    VisitInfo r;
    const int k = find_scan_by_visit(P, v);
    const DevScan &S = P.scans[k];
    unsigned unit, slot;
    split_unit(S, v, unit, slot);
    const unsigned R = (unsigned)P.restart;
    unsigned seg_q = 0, seg_r = unit; // unit = seg_q * R + seg_r
    if (R) divmod(unit, P.div_restart, seg_q, seg_r);
    const bool restart_here = unit == 0 || (R && seg_r == 0);
    unsigned long long blk, pred = 0;
    bool has_pred = true;
    int comp;
    if (S.comp < 0) {
        comp = P.slot_comp[slot];
        const unsigned bv = P.slot_v[slot], bh = P.slot_h[slot];
        const unsigned H = P.comp_h[comp], V = P.comp_v[comp], pw = P.comp_pw[comp];
        unsigned my, mx;
        divmod(unit, P.div_mcu_cols, my, mx);
        const unsigned long long off = P.comp_off[comp];
        blk = off + (unsigned long long)(my * V + bv) * pw + mx * H + bh;
        r.pred_back = (bh > 0 || bv > 0) ? 1 : (int)(S.bpu - H * V + 1);
        if (bh > 0) pred = blk - 1;
        else if (bv > 0) pred = off + (unsigned long long)(my * V + bv - 1) * pw + mx * H + (H - 1);
        else if (restart_here) has_pred = false;
        else { // last block of this component in the previous MCU
            const unsigned pmy = mx ? my : my - 1, pmx = mx ? mx - 1 : P.mcu_cols - 1;
            pred = off + (unsigned long long)(pmy * V + V - 1) * pw + pmx * H + (H - 1);
        }
    } else {
        comp = S.comp;
        r.pred_back = 1;
        const unsigned tw = P.comp_tw[comp], pw = P.comp_pw[comp];
        const unsigned long long off = P.comp_off[comp];
        unsigned by, bx;
        divmod(unit, P.div_tw[comp], by, bx);
        blk = off + (unsigned long long)by * pw + bx;
        if (restart_here) has_pred = false;
        else { // previous block of the raster grid
            const unsigned pby = bx ? by : by - 1, pbx = bx ? bx - 1 : tw - 1;
            pred = off + (unsigned long long)pby * pw + pbx;
        }
    }
    r.blk = coef_img + blk * 64;
    r.pred = has_pred ? coef_img + pred * 64 : nullptr;
    r.comp = comp;
    r.ss = S.ss;
    r.se = S.se;
    r.tbl = P.comp_tbl[comp];
    r.scan = k;
    r.seg_in_scan = seg_q;
    r.seg_local = S.seg_base + r.seg_in_scan;
    r.first_visit_of_seg = S.visit_base + (unsigned long long)r.seg_in_scan * R * S.bpu;
    r.last_of_seg = slot == S.bpu - 1 && (unit == S.n_units - 1 || (R && seg_r == R - 1));
    return r;
}

// get_code, writer.rs:455-470: size = bit length of |v|, bits = low `size` bits of (v - (v<0))
__device__ __forceinline__ void value_code(int v, int &size, uint32_t &bits) {
    const int a = v < 0 ? -v : v;
    int top; // index of the highest set bit, -1 for 0
    asm("bfind.u32 %0, %1;" : "=r"(top) : "r"(a));
    size = top + 1;
    uint32_t mask;
    asm("bmsk.clamp.b32 %0, 0, %1;" : "=r"(mask) : "r"(size));
    bits = (uint32_t)(v + (v >> 31)) & mask;
}

__device__ __forceinline__ unsigned nonzero16x2(unsigned x) {
    unsigned r;
    asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(0x00010001u));
    return r;
}

// Bit sink of one visit. Codes are packed MSB-first into 32-bit words. Slots are tiled by kSlotTile
// visits (one coding CTA): word j of visit g lives at slot_of(g)[j * kSlotTile], so the lanes of a warp
// write one contiguous line per word and a CTA's slots are one contiguous 56 KB region. (Word-major
// across the whole launch, planes n_visits words apart, made every CTA touch a dozen pages hundreds of
// MB apart and the coding kernel TLB-bound.) The visit is coded exactly once; where its bits belong in
// the stream is decided later, from the prefix sum of the totals, by place_bits_kernel.
__device__ __forceinline__ uint32_t *slot_of(uint32_t *slots, unsigned long long g) {
    return slots + (g / kSlotTile) * (kSlotTile * kSlotWords) + (g % kSlotTile);
}

struct BitSink {
    unsigned long long acc = 0;
    int n = 0;
    unsigned off = 0; // bytes written so far, scaled by the tile stride
    uint32_t *slot0;

    __device__ __forceinline__ BitSink(uint32_t *slots, unsigned long long g) : slot0(slot_of(slots, g)) {}
    __device__ __forceinline__ void put(uint32_t code, int len) { // len <= 31, n < 32 on entry
        acc = (acc << len) | code;
        n += len;
        if (n >= 32) {
            n -= 32;
            *reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(slot0) + off) = (uint32_t)(acc >> n);
            off += kSlotTile * 4;
        }
    }
    __device__ __forceinline__ unsigned finish() {
        if (n > 0) *reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(slot0) + off) = (uint32_t)(acc << (32 - n)); // left-aligned tail
        return off / (kSlotTile * 4) * 32 + n;
    }
};

// Table words as the host uploads them (api.cu): for the symbol (run << 4 | size), or the DC category
// `size`, (code_length + size) << 27 | code << size: OR-ing the value bits in gives the whole
// write_bits argument of huffman_encode_value (writer.rs:320-329). A symbol without a code is
// size << 27: only the value bits are written (release-build behaviour of the reference).
constexpr uint32_t kCodeBits = 0x07FFFFFFu;

constexpr int kEncThreads = kSlotTile;
constexpr int kStageStride = 72; // int16 per staged block: 128 B of coefficients + 16 B pad (128-bit rows stay conflict-free)

struct __align__(16) EncodeShared {
    int16_t coef[kEncThreads * kStageStride]; // the CTA's blocks, staged with coalesced loads
    unsigned long long ptr_or_mask[kEncThreads]; // first the block address | group range, then the non-zero mask
    uint32_t first[kEncThreads];  // the DC code of the visit (table-word format), 0 in AC scans
    uint32_t info[kEncThreads];   // se | first_ac << 8 | table << 16 | valid << 24
    uint32_t ac_tab[2 * 256];
    uint32_t dc_tab[2 * 16];
    uint32_t bin[64];
    uint16_t order[kEncThreads];
};

__device__ __forceinline__ const uint32_t *huff_for(const EntropyBuffers &b, unsigned long long img, int tbl, int cls) {
    return b.huff + (b.huff_per_image ? img * kHuffWordsPerImage : 0) + (size_t)(tbl * 2 + cls) * 256;
}

// Zero runs longer than 15 in front of a non-zero coefficient take one ZRL symbol per 16 zeros (writer.rs:369-373).
// `m` has bit k set for every non-zero coefficient k of the band [first_ac, se]. Returns the positions of the
// zeros that end such a group of 16: start of the run + 15, + 31, + 47.
__device__ __forceinline__ unsigned long long zrl_markers(unsigned long long m, int first_ac) {
    const unsigned long long z = m | (1ull << (first_ac - 1)); // the position in front of the band ends "run 0"
    unsigned long long s = z; // bit j: z has a set bit in [j - 15, j]
    s |= s << 1;
    s |= s << 2;
    s |= s << 4;
    s |= s << 8;
    unsigned long long need = m & ~(s << 1); // non-zeros with 16 or more zeros in front of them
    unsigned long long marks = 0;
    while (need) { // rare
        const int k = __ffsll((long long)need) - 1;
        need &= need - 1;
        const int j = 63 - __clzll((long long)(z & ((1ull << k) - 1ull))); // the set bit in front of k
        for (int t = j + 16; t < k; t += 16) marks |= 1ull << t;
    }
    return marks;
}

// write_ac_block restricted to the band (writer.rs:354-388) for one visit. The positions to code come as two
// bit-reversed masks: coefficient k of the half [BASE, BASE + 32) is bit 31 - (k - BASE), so the next one in
// zig-zag order is the highest set bit (one FLO, no bit reversal). One iteration per set bit: a non-zero
// coefficient, or a ZRL marker (zrl_markers): a zero 15 positions after the start of its run, for which the same
// arithmetic yields the symbol 0xF0 with no value bits (writer.rs:369-373), so the loop has no ZRL branch.
template <int BASE>
__device__ __forceinline__ void code_half(unsigned m, int &next, unsigned c_shared, const uint32_t *__restrict__ tab, BitSink &sink) {
#pragma unroll 1
    while (m) {
        int p;
        asm("bfind.u32 %0, %1;" : "=r"(p) : "r"(m));
        unsigned bit;
        asm("bmsk.clamp.b32 %0, %1, 1;" : "=r"(bit) : "r"(p));
        m ^= bit;
        const int k = BASE + 31 - p;
        int v;
        asm("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(c_shared + 2 * k));
        const int run = k - next; // next = first position of the current zero run; <= 15 thanks to the ZRL markers
        next = k + 1;
        int size;
        uint32_t bits;
        value_code(v, size, bits);
        const uint32_t e = (tab + run * 16)[size];
        sink.put((e & kCodeBits) | bits, (int)(e >> 27));
    }
}
__device__ __forceinline__ void code_nonzeros(unsigned lo_rev, unsigned hi_rev, int first_ac, int se, const int16_t *__restrict__ c,
                                              const uint32_t *__restrict__ tab, BitSink &sink) {
    int next = first_ac;
    const unsigned c_shared = (unsigned)__cvta_generic_to_shared(c);
    code_half<0>(lo_rev, next, c_shared, tab, sink);
    code_half<32>(hi_rev, next, c_shared, tab, sink);
    if (next <= se) { // the band ends in zeros: EOB (writer.rs:383-385)
        const uint32_t e = tab[0];
        sink.put(e & kCodeBits, (int)(e >> 27));
    }
}

// write_dc + write_ac_block for 256 consecutive visits per CTA, in three steps:
//  1. every thread locates its visit; each warp copies its 32 blocks into shared memory with
//     coalesced 128-bit loads (8 lanes per block);
//  2. every thread builds the 64-bit mask of non-zero coefficients of its block inside the scan's band
//     and the DC code; the CTA sorts its visits by their number of non-zeros (counting sort);
//  3. thread t codes the visit of rank t: the lanes of a warp then run nearly the same number of
//     iterations of the per-coefficient loop, which is where the time goes.
// FULL: every scan of the plan covers the whole block (baseline and sequential modes), so the band
// bookkeeping is constant.
template <bool FULL>
__global__ void __launch_bounds__(kEncThreads) encode_visits_kernel(const EntropyBuffers b, unsigned long long n_visits) {
    __shared__ EncodeShared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long cta_base = (unsigned long long)blockIdx.x * kEncThreads;
    const unsigned long long g = cta_base + tid;
    const bool valid = g < n_visits;
    const DevPlan &P = *b.plan;

    // AC tables of the CTA's first image into shared memory; a CTA that spans images with their own
    // (optimized) tables reads them from global memory instead
    unsigned long long img_first = 0, img_last = 0, tmp;
    if (b.huff_per_image) {
        split_visit(P, cta_base, img_first, tmp);
        split_visit(P, (cta_base + kEncThreads < n_visits ? cta_base + kEncThreads : n_visits) - 1, img_last, tmp);
    }
    const bool shared_tables = img_first == img_last;
    // the table words travel through registers: loaded here, stored to shared memory just before the first CTA
    // barrier, so their latency hides behind locating and staging
    static_assert(kEncThreads == 256, "two AC words and at most one DC word per thread");
    uint32_t tab_ac0 = 0, tab_ac1 = 0, tab_dc = 0;
    if (shared_tables) {
        const uint32_t *set = b.huff + img_first * kHuffWordsPerImage;
        tab_ac0 = __ldg(set + 256 + tid);       // table 0, AC
        tab_ac1 = __ldg(set + 512 + 256 + tid); // table 1, AC
        if (tid < 32) tab_dc = __ldg(set + (tid >> 4) * 512 + (tid & 15)); // DC categories 0..15 of both tables
    }
    if (tid < 64) sh.bin[tid] = 0;

    // ---- 1. locate, stage ----
    unsigned long long img = 0, v = 0;
    VisitInfo vi{};
    int w_lo = 1, w_hi = 0, first_ac = 1;
    unsigned long long where = 0;
    if (valid) {
        split_visit(P, g, img, v);
        vi = locate_visit(P, b.coef + img * P.blocks_per_image * 64, v);
        first_ac = FULL || vi.ss == 0 ? 1 : vi.ss;
        w_lo = FULL ? 0 : vi.ss >> 3;
        w_hi = FULL ? 7 : vi.se >> 3;
        where = (unsigned long long)vi.blk | (unsigned)w_lo | (unsigned)(w_hi << 3); // blocks are 128-byte aligned
    }
    if (!FULL && !__syncthreads_or(valid && vi.se > 0)) { // a CTA inside DC scans (progressive): one code per visit
        if (valid) {
            unsigned len = 0;
            // se == 0 is a DC scan (ss == 0) or the *empty* first AC band that 34..64 progressive scans produce
            // (64 / (scans - 1) == 1: band 0 = [1, 1), written as Ss=1, Se=0 -- encoder.rs:926-944): no bits at all
            if (vi.ss == 0) {
                const int prev = vi.pred ? (int)__ldg(vi.pred) : 0;
                int size;
                uint32_t bits;
                value_code((int)(int16_t)(__ldg(vi.blk) - prev), size, bits);
                const uint32_t e = __ldg(huff_for(b, img, vi.tbl, 0) + size) | bits;
                len = e >> 27;
                if (len) *slot_of(b.slots, g) = (e & kCodeBits) << (32 - len);
            }
            b.nbits[g] = len;
        }
        return;
    }
    sh.ptr_or_mask[tid] = where;
    __syncwarp();
    { // eight asynchronous 16-byte copies per lane, all in flight together
        const int w = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int bi = warp * 32 + i * 4 + (lane >> 3);
            const unsigned long long e = sh.ptr_or_mask[bi];
            const uint4 *src = reinterpret_cast<const uint4 *>(e & ~127ull);
            if (src != nullptr && (FULL || (w >= (int)(e & 7) && w <= (int)((e >> 3) & 7)))) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(sh.coef + bi * kStageStride + w * 8);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + w) : "memory");
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncwarp();

    // ---- 2. mask of non-zeros in [first_ac, se], DC code, sort key ----
    unsigned m_lo = 0, m_hi = 0; // bit-reversed: coefficient k is bit 31 - k % 32
    uint32_t first = 0;
    if (valid) {
        const uint4 *mine = reinterpret_cast<const uint4 *>(sh.coef + tid * kStageStride);
        if (FULL || vi.se > 0) {
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                if (!FULL && (w < w_lo || w > w_hi)) continue;
                const uint4 q = mine[w];
                // per 16-bit half min(x, 1) = (x != 0); dp2a weighs the two halves 1 and 2 and accumulates
                unsigned byte = __dp2a_lo(nonzero16x2(q.x), 0x0201u, 0u);
                byte = __dp2a_lo(nonzero16x2(q.y), 0x0804u, byte);
                byte = __dp2a_lo(nonzero16x2(q.z), 0x2010u, byte);
                byte = __dp2a_lo(nonzero16x2(q.w), 0x8040u, byte);
                if (w < 4) m_lo |= byte << (8 * w);
                else m_hi |= byte << (8 * (w - 4));
            }
            unsigned long long m = ((unsigned long long)m_hi << 32) | m_lo;
            m &= FULL ? ~1ull : (~0ull << first_ac) & (~0ull >> (63 - vi.se));
            m |= zrl_markers(m, first_ac);
            m_lo = __brev((unsigned)m); // coded from the highest bit down
            m_hi = __brev((unsigned)(m >> 32));
        }
    }
    const int nnz = __popc(m_lo) + __popc(m_hi); // <= 63
    if (shared_tables) {
        sh.ac_tab[tid] = tab_ac0;
        sh.ac_tab[256 + tid] = tab_ac1;
        if (tid < 32) sh.dc_tab[tid] = tab_dc;
    }
    __syncthreads();
    if (valid && (FULL || vi.ss == 0)) { // write_dc, writer.rs:342-352
        // The predecessor is the block of an earlier visit of this scan, a few visits back: staged by this CTA
        // (all staging is complete after the barrier) unless this visit is among the CTA's first.
        const int dc = sh.coef[tid * kStageStride];
        const int prev = !vi.pred ? 0 : (tid >= vi.pred_back ? (int)sh.coef[(tid - vi.pred_back) * kStageStride] : (int)__ldg(vi.pred));
        int size;
        uint32_t bits;
        value_code((int)(int16_t)(dc - prev), size, bits);
        first = (shared_tables ? sh.dc_tab[vi.tbl * 16 + size] : __ldg(huff_for(b, img, vi.tbl, 0) + size)) | bits;
    }
    const unsigned rank = atomicAdd(&sh.bin[nnz], 1u);
    sh.ptr_or_mask[tid] = ((unsigned long long)m_hi << 32) | m_lo;
    sh.first[tid] = first;
    sh.info[tid] = (unsigned)vi.se | ((unsigned)first_ac << 8) | ((unsigned)vi.tbl << 16) | (valid ? 1u << 24 : 0u);
    __syncthreads();
    { // exclusive prefix over the 64 bins, two per lane; every warp computes it for itself (no single-warp step)
        const unsigned a = sh.bin[2 * lane], c = sh.bin[2 * lane + 1];
        unsigned inc = a + c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        const unsigned even = inc - a - c, odd = inc - c; // first rank of bins 2*lane and 2*lane + 1
        const unsigned e = __shfl_sync(0xffffffffu, even, nnz >> 1), o = __shfl_sync(0xffffffffu, odd, nnz >> 1);
        sh.order[((nnz & 1) ? o : e) + rank] = (uint16_t)tid;
    }
    __syncthreads();

    // ---- 3. code the visit of rank tid ----
    const int src = sh.order[tid];
    const unsigned info = sh.info[src];
    if (!(info >> 24)) return;
    const unsigned long long mask = sh.ptr_or_mask[src];
    const uint32_t dc_code = sh.first[src];
    const int se = FULL ? 63 : info & 0xFF, fa = FULL ? 1 : (info >> 8) & 0xFF, tbl = (info >> 16) & 0xFF;
    BitSink sink(b.slots, cta_base + src);
    sink.put(dc_code & kCodeBits, (int)(dc_code >> 27));
    if (se > 0) {
        const int16_t *c = sh.coef + src * kStageStride;
        if (shared_tables) {
            code_nonzeros((unsigned)mask, (unsigned)(mask >> 32), fa, se, c, sh.ac_tab + tbl * 256, sink);
        } else {
            unsigned long long simg, sv;
            split_visit(P, cta_base + src, simg, sv);
            code_nonzeros((unsigned)mask, (unsigned)(mask >> 32), fa, se, c, huff_for(b, simg, tbl, 1), sink);
        }
    }
    b.nbits[cta_base + src] = sink.finish();
}

// lead of local segment `s`: what the reference writes between the previous segment's last byte
// and this segment's first: the file header (first segment of the image), the SOS of a later scan
// (writer.rs:424-452) or RSTn (encoder.rs:748-752: marker index = restarts % 8).
__device__ __forceinline__ unsigned lead_len(const EntropyBuffers &b, const DevPlan &P, unsigned long long img, int k,
                                             unsigned seg_in_scan) {
    if (P.scans[k].rst_base + seg_in_scan > 0) return 2; // RSTn (for a strip: also before its first segment)
    if (k > 0) return P.scans[k].sos_len;
    return b.hdr_len[b.huff_per_image ? img : 0];
}

__global__ void __launch_bounds__(256) segment_len_kernel(const EntropyBuffers b, unsigned long long n_segs) {
    const unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_segs) return;
    const DevPlan &P = *b.plan;
    const unsigned long long img = g / P.segs_per_image;
    const unsigned s = (unsigned)(g - img * P.segs_per_image);
    const int k = find_scan_by_seg(P, s);
    const DevScan &S = P.scans[k];
    const unsigned i = s - S.seg_base;
    const unsigned long long span = (unsigned long long)P.restart * S.bpu;
    const unsigned long long vf = S.visit_base + i * span;
    const unsigned long long vn = (i + 1 < S.n_segs) ? vf + span : S.visit_base + (unsigned long long)S.n_units * S.bpu;
    const unsigned long long vb = img * P.visits_per_image;
    const unsigned long long bits = b.bitpos[vb + vn] - b.bitpos[vb + vf];
    const unsigned tail = (s == P.segs_per_image - 1 && P.has_eoi) ? 2u : 0u; // EOI, encoder.rs:564
    b.seglen[g] = lead_len(b, P, img, k, i) + (unsigned)((bits + 7) >> 3) + tail;
}

// Unstuffed stream size as computed on the device; kernels that write the stream bail out (and the
// first one raises the overflow flag) when it exceeds the capacity the host provided.
__device__ __forceinline__ unsigned long long stream_bytes(const EntropyBuffers &b) { return b.segpos[b.n_segs_total]; }
__device__ __forceinline__ bool stream_fits(const EntropyBuffers &b) { return stream_bytes(b) <= b.ustream_cap; }

__global__ void __launch_bounds__(256) zero_ustream_kernel(const EntropyBuffers b, unsigned long long n_segs_total) {
    const unsigned long long bytes = b.segpos[n_segs_total];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        b.status[0] = bytes;
        if (bytes > b.ustream_cap) atomicOr(b.status + 2, 1ull);
    }
    if (bytes > b.ustream_cap) return;
    const unsigned long long n16 = (bytes + 15) >> 4, nm = ((bytes + 31) >> 5);
    uint4 *u = reinterpret_cast<uint4 *>(b.ustream);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
        u[i] = make_uint4(0, 0, 0, 0);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nm; i += stride) b.raw_mask[i] = 0;
}

__device__ __forceinline__ void put_raw(const EntropyBuffers &b, unsigned long long pos, uint8_t byte) {
    b.ustream[pos] = byte;
    atomicOr(b.raw_mask + (pos >> 5), 1u << (pos & 31));
}

__global__ void __launch_bounds__(128) segment_lead_kernel(const EntropyBuffers b, unsigned long long n_segs) {
    // one warp per segment; lanes stride over the lead bytes (headers can be long: ICC, EXIF)
    const unsigned long long g = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= n_segs || !stream_fits(b)) return;
    const DevPlan &P = *b.plan;
    const unsigned long long img = g / P.segs_per_image;
    const unsigned s = (unsigned)(g - img * P.segs_per_image);
    const int k = find_scan_by_seg(P, s);
    const DevScan &S = P.scans[k];
    const unsigned i = s - S.seg_base;
    const unsigned long long pos = b.segpos[g];
    if (S.rst_base + i > 0) {
        if (lane == 0) {
            put_raw(b, pos, 0xFF);
            put_raw(b, pos + 1, (uint8_t)(0xD0 + ((S.rst_base + i - 1) & 7)));
        }
    } else if (k > 0) {
        for (unsigned j = lane; j < S.sos_len; j += 32) put_raw(b, pos + j, P.blob[S.sos_off + j]);
    } else {
        const unsigned long long h = b.huff_per_image ? img : 0;
        const unsigned n = b.hdr_len[h];
        const uint8_t *src = b.hdr + h * b.hdr_stride;
        for (unsigned j = lane; j < n; j += 32) put_raw(b, pos + j, src[j]);
    }
    if (s == P.segs_per_image - 1 && P.has_eoi && lane == 0) {
        const unsigned long long end = b.segpos[g + 1];
        put_raw(b, end - 2, 0xFF);
        put_raw(b, end - 1, 0xD9);
    }
}

// Moves the already coded bits of visit g from its slot to their place in the unstuffed stream:
// a bit-granular copy (funnel shift by the start position modulo 32). The first and the last stream
// word of a visit are shared with its neighbours and are merged with atomicOr (the stream is
// zero-initialised); words in between are owned and stored. The last visit of a segment also
// writes the pad bits of finalize_bit_buffer (writer.rs:138-145: ones up to the byte boundary).
__global__ void __launch_bounds__(256) place_bits_kernel(const EntropyBuffers b, unsigned long long n_visits) {
    const unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_visits || !stream_fits(b)) return;
    const DevPlan &P = *b.plan;
    // The kernel is bound by the latency of its dependent loads, not by bytes or instructions: the first four slot
    // words are requested before anything else is known (every visit owns all kSlotWords rows of its tile, so the
    // loads are always in bounds; words beyond the visit's code are simply not used).
    const uint32_t *src = slot_of(b.slots, g);
    uint32_t pre[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) pre[q] = src[(size_t)q * kSlotTile];
    unsigned long long img, v;
    split_visit(P, g, img, v);
    unsigned nb = b.nbits[g];
    // segment bookkeeping only (no coefficient access)
    bool last_of_seg;
    unsigned long long data_byte, rel_bits;
    if (P.n_scans == 1 && P.restart == 0 && P.scans[0].rst_base == 0) {
        // one scan, no restarts (the interleaved baseline file): the image is one segment that starts with the header
        last_of_seg = v == P.visits_per_image - 1;
        if (nb == 0 && !last_of_seg) return;
        data_byte = b.segpos[img] + b.hdr_len[b.huff_per_image ? img : 0];
        rel_bits = b.bitpos[g] - b.bitpos[img * P.visits_per_image];
    } else {
        const int k = find_scan_by_visit(P, v);
        const DevScan &S = P.scans[k];
        unsigned unit, slot_in_unit;
        split_unit(S, v, unit, slot_in_unit);
        const unsigned R = (unsigned)P.restart;
        unsigned seg_in_scan = 0, seg_r = unit;
        if (R) divmod(unit, P.div_restart, seg_in_scan, seg_r);
        last_of_seg = slot_in_unit == S.bpu - 1 && (unit == S.n_units - 1 || (R && seg_r == R - 1));
        if (nb == 0 && !last_of_seg) return;
        const unsigned long long first_visit = S.visit_base + (unsigned long long)seg_in_scan * R * S.bpu;
        const unsigned long long seg = img * P.segs_per_image + S.seg_base + seg_in_scan;
        data_byte = b.segpos[seg] + lead_len(b, P, img, k, seg_in_scan);
        rel_bits = b.bitpos[g] - b.bitpos[img * P.visits_per_image + first_visit];
    }
    const unsigned long long bitpos = data_byte * 8 + rel_bits;

    uint32_t *dst = reinterpret_cast<uint32_t *>(b.ustream) + (bitpos >> 5);
    const unsigned sh = (unsigned)(bitpos & 31);
    const unsigned n_words = (nb + 31) >> 5;
    unsigned pad = 0;
    if (last_of_seg) {
        const unsigned end_bits = (unsigned)((rel_bits + nb) & 7);
        pad = end_bits ? 8 - end_bits : 0;
    }
    uint32_t carry = 0; // bits still to be written into the current destination word (left-aligned)
    bool first = true;
    auto emit = [&](unsigned j, uint32_t w) {
        const unsigned have = (j + 1 == n_words) ? nb - 32 * j : 32u; // valid bits in w (left-aligned)
        if (j + 1 == n_words && pad) { // append the pad ones behind the last code bits when they fit in this word
            if (have + pad <= 32) {
                w |= ((1u << pad) - 1u) << (32 - have - pad);
                pad = 0;
            }
        }
        const uint32_t out = carry | (sh ? (w >> sh) : w);
        const bool full = sh + have >= 32; // this destination word is completed by w
        const uint32_t be = __byte_perm(out, 0, 0x0123);
        if (first || !full) atomicOr(dst, be);
        else *dst = be;
        first = false;
        if (full) {
            ++dst;
            carry = sh ? (w << (32 - sh)) : 0u;
            if (j + 1 == n_words) { // bits of the last word that spilled into the next destination word
                const unsigned spill = sh + have - 32;
                if (spill) atomicOr(dst, __byte_perm(carry, 0, 0x0123));
            }
        }
    };
#pragma unroll
    for (unsigned j = 0; j < 4; ++j)
        if (j < n_words) emit(j, pre[j]);
    for (unsigned j = 4; j < n_words; ++j) emit(j, src[(size_t)j * kSlotTile]);
    if (pad) { // pad bits that did not fit next to the last code word (or a visit without bits)
        const unsigned long long p = bitpos + nb;
        uint32_t *d2 = reinterpret_cast<uint32_t *>(b.ustream) + (p >> 5);
        const unsigned s2 = (unsigned)(p & 31); // pad never crosses a byte, hence never a word
        atomicOr(d2, __byte_perm(((1u << pad) - 1u) << (32 - s2 - pad), 0, 0x0123));
    }
}

// ---- 0xFF stuffing (writer.rs:156-167) as count / scan / scatter ---------------------------------
// A thread owns 16 consecutive bytes of the unstuffed stream and the 16 raw_mask bits beside them.
// 0x80 in every byte of w that equals 0xFF, exact (no carries cross bytes): a byte of ~w is zero iff neither its
// low seven bits (sum with 0x7F stays below 0x80) nor its top bit are set.
__device__ __forceinline__ uint32_t ff_flags(uint32_t w) {
    const uint32_t x = ~w;
    const uint32_t t = ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x;
    return ~t & 0x80808080u;
}
// flags at bits 7, 15, 23, 31 -> bits 0..3 (the four products land on distinct bits 21..24)
__device__ __forceinline__ unsigned flags_to_nibble(uint32_t z) { return (((z >> 7) * 0x00204081u) >> 21) & 0xFu; }

__device__ __forceinline__ unsigned ff_bits16(const uint4 d, unsigned raw16, unsigned valid) {
    unsigned m = flags_to_nibble(ff_flags(d.x)) | flags_to_nibble(ff_flags(d.y)) << 4 | flags_to_nibble(ff_flags(d.z)) << 8 |
                 flags_to_nibble(ff_flags(d.w)) << 12;
    m &= ~raw16;
    if (valid < 16) m &= (1u << valid) - 1u;
    return m;
}
// number of data 0xFF bytes among the 16 (same as popc(ff_bits16), cheaper when nothing is raw)
__device__ __forceinline__ unsigned ff_count16(const uint4 d, unsigned raw16, unsigned valid) {
    if (raw16 == 0 && valid >= 16) return __popc(ff_flags(d.x)) + __popc(ff_flags(d.y)) + __popc(ff_flags(d.z)) + __popc(ff_flags(d.w));
    return __popc(ff_bits16(d, raw16, valid));
}

__global__ void __launch_bounds__(256) count_ff_kernel(const EntropyBuffers b) {
    __shared__ unsigned warp_sums[8];
    const unsigned long long bytes = stream_fits(b) ? stream_bytes(b) : 0ull;
    const unsigned long long base = (unsigned long long)blockIdx.x * kStuffChunk + (unsigned long long)threadIdx.x * 16;
    unsigned cnt = 0;
    if (base < bytes) {
        const uint4 d = *reinterpret_cast<const uint4 *>(b.ustream + base);
        const unsigned raw = (b.raw_mask[base >> 5] >> (base & 31)) & 0xFFFFu;
        const unsigned valid = bytes - base < 16 ? (unsigned)(bytes - base) : 16u;
        cnt = ff_count16(d, raw, valid);
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int i = 0; i < 8; ++i) t += warp_sums[i];
        b.ffcount[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) stuff_scatter_kernel(const EntropyBuffers b, unsigned long long n_chunks_cap) {
    if (!stream_fits(b)) return;
    const unsigned long long bytes = stream_bytes(b);
    {
        const unsigned long long ff_total = b.ffpos[n_chunks_cap];
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            b.status[1] = ff_total;
            if (bytes + ff_total > b.out_cap) atomicOr(b.status + 2, 2ull);
        }
        if (bytes + ff_total > b.out_cap) return;
    }
    if ((unsigned long long)blockIdx.x * kStuffChunk >= bytes) return;
    // The chunk's output is first laid out in shared memory at the same offset modulo 16 as its place
    // in `out`, then copied with aligned 128-bit stores (single bytes only at the two ends).
    __shared__ unsigned warp_excl[9];
    __shared__ __align__(16) uint8_t stage[2 * kStuffChunk + 32];
    const unsigned long long base = (unsigned long long)blockIdx.x * kStuffChunk + (unsigned long long)threadIdx.x * 16;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 d = make_uint4(0, 0, 0, 0);
    unsigned m = 0, valid = 0;
    if (base < bytes) {
        d = *reinterpret_cast<const uint4 *>(b.ustream + base);
        const unsigned raw = (b.raw_mask[base >> 5] >> (base & 31)) & 0xFFFFu;
        valid = bytes - base < 16 ? (unsigned)(bytes - base) : 16u;
        m = ff_bits16(d, raw, valid);
    }
    const unsigned cnt = __popc(m);
    unsigned inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_excl[warp + 1] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        warp_excl[0] = 0;
        for (int i = 1; i <= 8; ++i) warp_excl[i] += warp_excl[i - 1];
    }
    __syncthreads();
    const unsigned long long chunk_base = (unsigned long long)blockIdx.x * kStuffChunk;
    const unsigned long long out0 = chunk_base + b.ffpos[blockIdx.x]; // first output byte of this chunk
    const unsigned pad = (unsigned)(out0 & 15);
    const unsigned chunk_in = bytes - chunk_base < kStuffChunk ? (unsigned)(bytes - chunk_base) : (unsigned)kStuffChunk;
    const unsigned total = chunk_in + warp_excl[8];
    if (valid) {
        unsigned o = pad + threadIdx.x * 16 + warp_excl[warp] + (inc - cnt);
        const uint32_t w[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if ((unsigned)i < valid) {
                stage[o++] = (uint8_t)(w[i >> 2] >> ((i & 3) * 8));
                if (m & (1u << i)) stage[o++] = 0x00;
            }
        }
    }
    __syncthreads();
    uint8_t *dst = b.out + (out0 - pad); // 16-byte aligned (b.out comes from cudaMalloc)
    const unsigned end = pad + total;
    for (unsigned o = threadIdx.x * 16; o < end; o += 256 * 16) {
        if (o >= pad && o + 16 <= end) {
            *reinterpret_cast<uint4 *>(dst + o) = *reinterpret_cast<const uint4 *>(stage + o);
        } else {
            for (unsigned k = (o < pad ? pad : o); k < o + 16 && k < end; ++k) dst[k] = stage[k];
        }
    }
}

// data 0xFF bytes in [chunk start, pos) of the unstuffed stream, counted by one warp with 128-bit loads
__device__ __forceinline__ unsigned ff_before_in_chunk(const EntropyBuffers &b, unsigned long long pos, int lane) {
    const unsigned long long start = pos / kStuffChunk * kStuffChunk;
    unsigned ff = 0;
    for (unsigned long long i = start + (unsigned long long)lane * 16; i < pos; i += 32 * 16) {
        const uint4 d = *reinterpret_cast<const uint4 *>(b.ustream + i);
        const unsigned raw = (b.raw_mask[i >> 5] >> (i & 31)) & 0xFFFFu;
        const unsigned valid = pos - i < 16 ? (unsigned)(pos - i) : 16u;
        ff += ff_count16(d, raw, valid);
    }
    return __reduce_add_sync(0xffffffffu, ff);
}

// byte offset of every file in `out`: position of the image's first segment plus the data 0xFF
// bytes that precede it
__global__ void __launch_bounds__(128) file_offsets_kernel(const EntropyBuffers b, unsigned n_images) {
    if (!stream_fits(b)) return;
    const unsigned long long bytes = stream_bytes(b);
    // one warp per file boundary; lanes stride over the (< kStuffChunk) bytes between the chunk start and it
    const unsigned img = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (img > n_images) return;
    const DevPlan &P = *b.plan;
    const unsigned long long pos = img == n_images ? bytes : b.segpos[(unsigned long long)img * P.segs_per_image];
    const unsigned long long chunk = pos / kStuffChunk;
    const unsigned ff = ff_before_in_chunk(b, pos, lane);
    if (lane == 0) b.file_off[img] = pos + b.ffpos[chunk] + ff;
}

// final byte offset of the first segment of every scan of image 0 (+ the end): the pieces of a strip
__global__ void __launch_bounds__(128) scan_offsets_kernel(const EntropyBuffers b, unsigned long long *offs) {
    if (!stream_fits(b)) return;
    const unsigned long long bytes = stream_bytes(b);
    const unsigned k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const DevPlan &P = *b.plan;
    if (k > (unsigned)P.n_scans) return;
    const unsigned long long pos = k == (unsigned)P.n_scans ? bytes : b.segpos[P.scans[k].seg_base];
    const unsigned long long chunk = pos / kStuffChunk;
    const unsigned ff = ff_before_in_chunk(b, pos, lane);
    if (lane == 0) offs[k] = pos + b.ffpos[chunk] + ff;
}

// ---- optimized-table histogram (encoder.rs:1086-1200) --------------------------------------------
// One thread per block of each component's *true* grid. DC category of the chained difference with
// no restart resets (Q17); AC run/size symbols per progressive band (runs restart per band), ZRL
// for runs > 15, EOB when a band ends in zeros. Bins: [image][table][dc|ac][257].
__global__ void __launch_bounds__(256) histogram_kernel(const DevPlan *plan, const int16_t *coef, unsigned long long n_blocks_total,
                                                        unsigned long long blocks_true_per_image, uint32_t *hist,
                                                        int bands, int per_band) {
    __shared__ unsigned sh[2 * 2 * 257];
    for (int i = threadIdx.x; i < 2 * 2 * 257; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const DevPlan &P = *plan;
    const unsigned long long g0 = (unsigned long long)blockIdx.x * blockDim.x;
    const unsigned long long g = g0 + threadIdx.x;
    const unsigned long long img_cta = g0 / blocks_true_per_image; // CTAs never straddle images (grid is per image)
    if (g < n_blocks_total) {
        unsigned long long r = g - img_cta * blocks_true_per_image;
        int comp = 0;
        for (; comp < P.ncomp - 1; ++comp) {
            const unsigned long long nb = (unsigned long long)P.comp_tw[comp] * (P.scans[comp].n_units / P.comp_tw[comp]);
            if (r < nb) break;
            r -= nb;
        }
        const unsigned tw = P.comp_tw[comp], pw = P.comp_pw[comp];
        const unsigned by = (unsigned)(r / tw), bx = (unsigned)(r - (unsigned long long)by * tw);
        const int16_t *img_coef = coef + img_cta * P.blocks_per_image * 64;
        const int16_t *blk = img_coef + (P.comp_off[comp] + (unsigned long long)by * pw + bx) * 64;
        int prev = 0;
        if (r > 0) {
            const unsigned long long pr = r - 1;
            const unsigned pby = (unsigned)(pr / tw), pbx = (unsigned)(pr - (unsigned long long)pby * tw);
            prev = img_coef[(P.comp_off[comp] + (unsigned long long)pby * pw + pbx) * 64];
        }
        unsigned *h = sh + P.comp_tbl[comp] * 2 * 257;
        const int diff = (int)(int16_t)(blk[0] - prev);
        atomicAdd(h + (32 - __clz(diff < 0 ? -diff : diff)), 1u);
        unsigned *ha = h + 257;
        const uint4 *src = reinterpret_cast<const uint4 *>(blk);
        int run = 0, band = 0, band_end = bands == 1 ? 64 : per_band; // band b covers [max(b*per,1), (b+1)*per), last to 64
        for (int w = 0; w < 8; ++w) {
            const uint4 q = __ldg(src + w);
            const uint32_t words[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = w * 8 + j;
                if (k == 0) continue;
                if (k == band_end) { // band boundary: close the previous band
                    if (run > 0) atomicAdd(ha, 1u);
                    run = 0;
                    ++band;
                    band_end = band == bands - 1 ? 64 : (band + 1) * per_band;
                }
                const int c = (int)(int16_t)(words[j >> 1] >> ((j & 1) * 16));
                if (c == 0) {
                    ++run;
                } else {
                    for (; run > 15; run -= 16) atomicAdd(ha + 0xF0, 1u);
                    atomicAdd(ha + ((run << 4) | (32 - __clz(c < 0 ? -c : c))), 1u);
                    run = 0;
                }
            }
        }
        if (run > 0) atomicAdd(ha, 1u);
    }
    __syncthreads();
    uint32_t *dst = hist + img_cta * (2 * 2 * 257);
    for (int i = threadIdx.x; i < 2 * 2 * 257; i += blockDim.x)
        if (sh[i]) atomicAdd(dst + i, sh[i]);
}

} // namespace

static inline unsigned grid_for(unsigned long long n, unsigned block) { return (unsigned)((n + block - 1) / block); }

cudaError_t launch_histogram(const DevPlan *plan, const DevPlan &hp, const int16_t *coef, uint32_t n_images, uint32_t *hist,
                             cudaStream_t stream) {
    // sequential / progressive plans only: scans 0..ncomp-1 are one per component over its true grid
    unsigned long long per_image = 0;
    for (int c = 0; c < hp.ncomp; ++c) per_image += hp.scans[c].n_units;
    const int bands = hp.n_scans > hp.ncomp ? hp.n_scans / hp.ncomp - 1 : 1;
    const int per_band = bands > 1 ? 64 / bands : 64;
    const unsigned ctas_per_image = grid_for(per_image, 256);
    // one launch per image keeps CTAs from straddling images; images in a batch are few in optimized mode
    for (uint32_t i = 0; i < n_images; ++i) {
        histogram_kernel<<<ctas_per_image, 256, 0, stream>>>(plan, coef + (size_t)i * hp.blocks_per_image * 64, per_image, per_image,
                                                            hist + (size_t)i * (2 * 2 * 257), bands, per_band);
    }
    return cudaGetLastError();
}

cudaError_t launch_symbol_sizes(const EntropyBuffers &b, const DevPlan &hp, uint32_t n, cudaStream_t s) {
    const unsigned long long nv = hp.visits_per_image * n;
    // progressive plans end with an AC band; in every other mode all scans cover the whole block
    const bool full = hp.scans[hp.n_scans - 1].ss == 0 && hp.scans[hp.n_scans - 1].se == 63;
    if (full) encode_visits_kernel<true><<<grid_for(nv, kEncThreads), kEncThreads, 0, s>>>(b, nv);
    else encode_visits_kernel<false><<<grid_for(nv, kEncThreads), kEncThreads, 0, s>>>(b, nv);
    return cudaGetLastError();
}
cudaError_t launch_segment_lengths(const EntropyBuffers &b, const DevPlan &hp, uint32_t n, cudaStream_t s) {
    const unsigned long long ns = (unsigned long long)hp.segs_per_image * n;
    segment_len_kernel<<<grid_for(ns, 256), 256, 0, s>>>(b, ns);
    return cudaGetLastError();
}
cudaError_t launch_zero_ustream(const EntropyBuffers &b, uint64_t n_segs_total, cudaStream_t s) {
    zero_ustream_kernel<<<148 * 8, 256, 0, s>>>(b, n_segs_total);
    return cudaGetLastError();
}
cudaError_t launch_segment_leads(const EntropyBuffers &b, const DevPlan &hp, uint32_t n, cudaStream_t s) {
    const unsigned long long ns = (unsigned long long)hp.segs_per_image * n;
    segment_lead_kernel<<<grid_for(ns * 32, 128), 128, 0, s>>>(b, ns);
    return cudaGetLastError();
}
cudaError_t launch_emit_bits(const EntropyBuffers &b, const DevPlan &hp, uint32_t n, cudaStream_t s) {
    const unsigned long long nv = hp.visits_per_image * n;
    place_bits_kernel<<<grid_for(nv, 256), 256, 0, s>>>(b, nv);
    return cudaGetLastError();
}
cudaError_t launch_count_ff(const EntropyBuffers &b, cudaStream_t s) {
    count_ff_kernel<<<grid_for(b.ustream_cap, kStuffChunk), 256, 0, s>>>(b);
    return cudaGetLastError();
}
cudaError_t launch_stuff_scatter(const EntropyBuffers &b, cudaStream_t s) {
    const unsigned long long n_chunks_cap = (b.ustream_cap + kStuffChunk - 1) / kStuffChunk;
    stuff_scatter_kernel<<<grid_for(b.ustream_cap, kStuffChunk), 256, 0, s>>>(b, n_chunks_cap);
    return cudaGetLastError();
}
cudaError_t launch_scan_offsets(const EntropyBuffers &b, const DevPlan &hp, unsigned long long *offs, cudaStream_t s) {
    scan_offsets_kernel<<<grid_for((hp.n_scans + 1ull) * 32, 128), 128, 0, s>>>(b, offs);
    return cudaGetLastError();
}
cudaError_t launch_file_offsets(const EntropyBuffers &b, const DevPlan &, uint32_t n, cudaStream_t s) {
    file_offsets_kernel<<<grid_for((n + 1ull) * 32, 128), 128, 0, s>>>(b, n);
    return cudaGetLastError();
}

} // namespace jpgb
