// Micro-benchmarks of the sm_100a integer pipes used by the field arithmetic (measure, don't guess): per-SM issue rate of
// IMAD / IMAD.WIDE / IMAD.HI / IADD3 / VIADDMNMX / LOP3 / SHF and of 1:1 mixes. Build: make -C scripts/ubench; run on the GPU box.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2048
template <int OP>
__global__ void __launch_bounds__(256) k_pipe(uint32_t* out, uint32_t seed) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 17 + i;
    uint32_t b = seed | 1, c = seed * 3 + 1;
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 1) {
                uint64_t t;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[i]), "r"(b));
                a[i] = (uint32_t)t ^ (uint32_t)(t >> 32);
            }
            if (OP == 2) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 4) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; min.u32 %0, t, %0; }" : "+r"(a[i]) : "r"(c));
            if (OP == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 6) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b));
            if (OP == 7) {  // 1:1 IMAD + IADD
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            }
            if (OP == 8) {  // Montgomery product (4 instr) 
                uint32_t lo, hi, m, r;
                asm volatile("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a[i]), "r"(b));
                m = lo * 0x7effffffu;
                asm volatile("{ .reg .u32 l2;\n\tmad.lo.cc.u32 l2, %1, %2, %3;\n\tmadc.hi.u32 %0, %1, %2, %4; }" : "=r"(r) : "r"(m), "r"(0x7f000001u), "r"(lo), "r"(hi));
                uint32_t r2 = r - 0x7f000001u;
                a[i] = r2 < r ? r2 : r;
            }
            if (OP == 9) {  // Shoup product by a fixed multiplier w with w' = floor(w * 2^32 / P): a*w - hi(a*w')*P in [0, 2P)
                const uint32_t P = 0x7f000001u;
                uint32_t q = __umulhi(a[i], c);   // c plays w'
                uint32_t r = a[i] * b - q * P;    // b plays w
                uint32_t r2 = r - P;
                a[i] = r2 < r ? r2 : r;
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Dependent-chain latency (cycles per link), one warp.
template <int OP>
__global__ void k_lat(uint32_t* out, long long* cyc, uint32_t seed) {
    uint32_t x = seed + threadIdx.x, b = seed | 1;
    const uint32_t P = 0x7f000001u;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (OP == 0) x = __shfl_xor_sync(0xffffffffu, x, 1);
            if (OP == 1) { uint32_t s = x + b, s2 = s - P; x = s2 < s ? s2 : s; }            // fadd
            if (OP == 2) {                                                                    // fmul
                uint64_t t = (uint64_t)x * b; uint32_t m = (uint32_t)t * 0x7effffffu, r;
                asm("{ .reg .u32 l2;\n\tmad.lo.cc.u32 l2, %1, %2, %3;\n\tmadc.hi.u32 %0, %1, %2, %4; }" : "=r"(r) : "r"(m), "r"(P), "r"((uint32_t)t), "r"((uint32_t)(t >> 32)));
                uint32_t r2 = r - P; x = r2 < r ? r2 : r;
            }
            if (OP == 3) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x) : "r"(b));     // IMAD
            if (OP == 4) { uint64_t t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(x), "r"(b)); x = (uint32_t)(t >> 32) | 1; }  // IMAD.WIDE (+LOP)
            if (OP == 5) asm volatile("mad.hi.u32 %0, %0, %1, %1;" : "+r"(x) : "r"(b));     // IMAD.HI
            if (OP == 6) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(b));            // IADD
            if (OP == 7) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; min.u32 %0, t, %0; }" : "+r"(x) : "r"(b));  // VIADDMNMX
            if (OP == 8) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 15);
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
static void lat(const char* name, uint32_t* d) {
    long long* cyc;
    cudaMallocManaged(&cyc, 8);
    k_lat<OP><<<1, 32>>>(d, cyc, 12345);
    cudaDeviceSynchronize();
    k_lat<OP><<<1, 32>>>(d, cyc, 12345);
    cudaDeviceSynchronize();
    printf("latency %-24s %6.1f cycles/link\n", name, (double)*cyc / (256 * 8));
    cudaFree(cyc);
}

template <int OP>
static void run(const char* name, int ops_per_slot, uint32_t* d) {
    int dev_sms;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    dim3 grid(dev_sms * 8), block(256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_pipe<OP><<<grid, block>>>(d, 12345);
    cudaDeviceSynchronize();
    float best = 1e9;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k_pipe<OP><<<grid, block>>>(d, 12345 + rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double lane_ops = (double)grid.x * 256 * ITER * 8 * ops_per_slot;
    double per_sm_clk = lane_ops / (best * 1e-3) / dev_sms / (clk * 1e3);
    printf("%-28s %8.3f ms  %7.1f lane-ops/clk/SM (at nominal %d MHz)  %.2f Tops/s\n", name, best, per_sm_clk, clk / 1000, lane_ops / (best * 1e-3) / 1e12);
}

int main() {
    uint32_t* d;
    cudaMalloc(&d, 148 * 8 * 256 * 4 * 2);
    run<0>("IMAD (mad.lo)", 1, d);
    run<1>("IMAD.WIDE (+LOP)", 1, d);
    run<2>("IMAD.HI (mad.hi)", 1, d);
    run<3>("IADD3", 1, d);
    run<4>("IADD3+VIMNMX (fadd)", 1, d);
    run<5>("LOP3", 1, d);
    run<6>("SHF", 1, d);
    run<7>("IMAD+IADD3 1:1 (pairs)", 2, d);
    run<9>("Shoup product (fixed w)", 1, d);
    run<8>("Montgomery fmul", 1, d);
    lat<0>("SHFL.BFLY", d);
    lat<8>("SHFL.IDX", d);
    lat<1>("fadd (IADD+VIADDMNMX)", d);
    lat<2>("fmul (4 instr)", d);
    lat<3>("IMAD", d);
    lat<4>("IMAD.WIDE(+LOP)", d);
    lat<5>("IMAD.HI", d);
    lat<6>("IADD3", d);
    lat<7>("VIADDMNMX", d);
    return 0;
}
