// Integer-pipe throughput probe for B200 (sm_100a): which ops are cheap for the colour/DCT/quant kernel.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int OP>
__global__ void k(unsigned *out, unsigned a0, unsigned b0) {
    unsigned x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a0 + threadIdx.x * 17 + i;
    unsigned b = b0 + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) x[i] = x[i] * b + 0x1234;                       // IMAD
            if (OP == 1) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(b), "r"(0x01020304u));
            if (OP == 2) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(b), "r"(0x01020304u));
            if (OP == 3) x[i] = __byte_perm(x[i], b, 0x7632);            // PRMT
            if (OP == 4) asm volatile("shr.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(b & 1));     // SHF
            if (OP == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(x[i]) : "r"(b), "r"(0x55aa55aau)); // LOP3
            if (OP == 6) asm volatile("add.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));          // IADD
            if (OP == 7) x[i] = ((int)x[i] < 0) ? b : x[i] + 1;           // ISETP+SEL-ish
            if (OP == 8) asm volatile("dp2a.hi.s32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(b), "r"(0x01020304u));
            if (OP == 9) x[i] = __funnelshift_r(x[i], b, 15);             // SHF.R funnel
            if (OP == 10) x[i] = abs((int)x[i]) + 1;                      // IABS
            if (OP == 12) { asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(b), "r"(0x01020304u)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(x[i]) : "r"(b), "r"(0x55aa55aau)); }
            if (OP == 13) { asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(b), "r"(0x01020304u)); x[i] = x[i] * b + 0x1234; }
            if (OP == 14) { asm volatile("add.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(b)); x[i] = x[i] * b + 0x1234; }
            if (OP == 15) { asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(x[i]) : "r"(b)); x[i] = x[i] * b + 0x1234; }
            if (OP == 16) { asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(x[i]) : "r"(b)); asm volatile("add.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(b)); }
            if (OP == 17) { asm volatile("mad.lo.s32 %0, %0, 1, %1;" : "+r"(x[i]) : "r"(b)); asm volatile("add.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(b)); }
            if (OP == 11) { x[i] = x[i] * b + 0x1234; x[i] = (x[i] >> 3) ^ b; } // IMAD + LOP/SHF pair (dual pipe)
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char *name, unsigned *d, int sms, double ops_per_iter = 1.0) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int blocks = sms * 8, threads = 256;
    k<OP><<<blocks, threads>>>(d, 1, 3);
    cudaEventRecord(a);
    k<OP><<<blocks, threads>>>(d, 1, 3);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ops = (double)blocks * threads * ITERS * ILP * ops_per_iter;
    printf("%-28s %8.3f ms  %8.2f Tops/s  (%.1f lane-ops/clk/SM at %d MHz nominal)\n", name, ms, ops / ms / 1e9,
           ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned *d;
    cudaMalloc(&d, sms * 8 * 256 * 4);
    printf("SMs: %d\n", sms);
    run<0>("IMAD", d, sms);
    run<1>("IDP.4A (dp4a u32.u32)", d, sms);
    run<2>("IDP.2A (dp2a.lo u32.u32)", d, sms);
    run<8>("IDP.2A (dp2a.hi s32.u32)", d, sms);
    run<3>("PRMT", d, sms);
    run<4>("SHF/SHR", d, sms);
    run<5>("LOP3", d, sms);
    run<6>("IADD3", d, sms);
    run<7>("ISETP+SEL", d, sms, 2.0);
    run<9>("SHF funnel", d, sms);
    run<10>("IABS+IADD", d, sms, 2.0);
    run<11>("IMAD + SHF + LOP3 mix", d, sms, 3.0);
    run<12>("IDP.2A + LOP3", d, sms, 2.0);
    run<13>("IDP.2A + IMAD", d, sms, 2.0);
    run<14>("IADD + IMAD", d, sms, 2.0);
    run<15>("PRMT + IMAD", d, sms, 2.0);
    run<16>("PRMT + IADD", d, sms, 2.0);
    run<17>("mad(x,1,b) + IADD", d, sms, 2.0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
