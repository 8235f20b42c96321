// Does running two permutations per thread half a permutation apart (one in its multiplier-heavy full rounds while the other is
// in its adder-heavy partial rounds) raise the issue rate over one permutation per thread? Also: odd warps delayed by D cycles.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../plonky3-recursion_b200/csrc/poseidon2.cuh"
using namespace p3r;

template <class F>
__device__ __forceinline__ void full_round(uint32_t* s, const Poseidon2Consts& k, int r) {
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = sbox<F>(fadd<F>(s[i], k.ext_rc[16 * r + i]));
    external_linear<F>(s, k.zero);
}
// 4 full rounds of x (rounds fr..fr+3) interleaved with 10 partial rounds of y (rounds pr..pr+9).
template <class F>
__device__ __forceinline__ void pair_phase(uint32_t* x, int fr, uint32_t* y, int pr, const Poseidon2Consts& k) {
#pragma unroll 1
    for (int j = 0; j < 2; j++) {
        full_round<F>(x, k, fr + 2 * j);
        internal_round_fast<F>(y, k.int_rc[pr + 5 * j]);
        internal_round_fast<F>(y, k.int_rc[pr + 5 * j + 1]);
        full_round<F>(x, k, fr + 2 * j + 1);
        internal_round_fast<F>(y, k.int_rc[pr + 5 * j + 2]);
        internal_round_fast<F>(y, k.int_rc[pr + 5 * j + 3]);
        internal_round_fast<F>(y, k.int_rc[pr + 5 * j + 4]);
    }
}

template <class F, int BS, int MINB>
__global__ void __launch_bounds__(BS, MINB) k_dual(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, int reps) {
    uint32_t i = blockIdx.x * BS + threadIdx.x;
    if (2 * i >= n) return;
    const Poseidon2Consts& k = c_p2[FieldId<F>::value];
    uint32_t a[16], b[16];
#pragma unroll
    for (int q = 0; q < 16; q++) a[q] = in[(size_t)q * n + 2 * i], b[q] = in[(size_t)q * n + 2 * i + 1];
    // prologue: b runs its first half alone
    external_linear<F>(b, k.zero);
#pragma unroll 1
    for (int r = 0; r < 4; r++) full_round<F>(b, k, r);
#pragma unroll 1
    for (int r = 0; r < 10; r++) internal_round_fast<F>(b, k.int_rc[r]);
#pragma unroll 1
    for (int it = 0; it < 2 * reps - 1; it++) {
        external_linear<F>(a, k.zero);
        pair_phase<F>(a, 0, b, 10, k);   // a: full 0-3, b: partial 10-19
        pair_phase<F>(b, 4, a, 0, k);    // b: full 4-7 (done), a: partial 0-9
#pragma unroll
        for (int q = 0; q < 16; q++) {   // roles swap: the finished state starts its next permutation as `a`
            uint32_t t = a[q];
            a[q] = b[q];
            b[q] = t;
        }
    }
    // epilogue: b (half done) finishes alone
#pragma unroll 1
    for (int r = 10; r < 20; r++) internal_round_fast<F>(b, k.int_rc[r]);
#pragma unroll 1
    for (int r = 4; r < 8; r++) full_round<F>(b, k, r);
    uint4* o = reinterpret_cast<uint4*>(out + (size_t)i * 8);
    o[0] = make_uint4(a[0] ^ b[0], a[1] ^ b[1], a[2] ^ b[2], a[3] ^ b[3]);
    o[1] = make_uint4(a[4] ^ b[4], a[5] ^ b[5], a[6] ^ b[6], a[7] ^ b[7]);
}

template <class F, int BS, int MINB>
__global__ void __launch_bounds__(BS, MINB) k_single(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, int reps, int delay) {
    uint32_t i = blockIdx.x * BS + threadIdx.x;
    if (i >= n) return;
    if (delay && ((threadIdx.x >> 5) & 1)) {
        long long t0 = clock64();
        while (clock64() - t0 < delay) {}
    }
    uint32_t st[16];
#pragma unroll
    for (int q = 0; q < 16; q++) st[q] = in[(size_t)q * n + i];
    for (int r = 0; r < reps; r++) poseidon2_permute<F>(st);
    uint4* o = reinterpret_cast<uint4*>(out + (size_t)i * 8);
    o[0] = make_uint4(st[0], st[1], st[2], st[3]);
    o[1] = make_uint4(st[4], st[5], st[6], st[7]);
}

template <class K>
static void timeit(const char* name, double perms, K launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e9;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    printf("%-44s %8.3f ms  %6.2f perms/ns  %s\n", name, best, perms / (best * 1e6), err ? cudaGetErrorString(err) : "");
}

int main() {
    Poseidon2Consts h[2];
    for (int f = 0; f < 2; f++) {
        uint32_t x = 12345 + f;
        uint32_t* w = reinterpret_cast<uint32_t*>(&h[f]);
        for (size_t i = 0; i < sizeof(Poseidon2Consts) / 4; i++) { x = x * 1664525u + 1013904223u; w[i] = x % 0x78000001u; }
        h[f].zero = 0;
        h[f].fast_diag = 1;
    }
    cudaMemcpyToSymbol(c_p2, h, sizeof(h));
    const uint32_t N = 1u << 20;
    const int R = 8;
    uint32_t *in, *out;
    cudaMalloc(&in, (size_t)N * 16 * 4);
    cudaMalloc(&out, (size_t)N * 8 * 4);
    cudaMemset(in, 1, (size_t)N * 16 * 4);
    for (int delay : {0, 350, 700, 1400, 2100})  {
        char nm[64];
        snprintf(nm, sizeof nm, "koala single, odd warps +%d cycles", delay);
        timeit(nm, (double)N * R, [&] { k_single<KoalaBear, 128, 1><<<N / 128, 128>>>(in, out, N, R, delay); });
    }
    timeit("koala dual bs128", (double)N * R, [&] { k_dual<KoalaBear, 128, 1><<<N / 2 / 128, 128>>>(in, out, N, R); });
    timeit("koala dual bs128 minb6", (double)N * R, [&] { k_dual<KoalaBear, 128, 6><<<N / 2 / 128, 128>>>(in, out, N, R); });
    timeit("koala dual bs64", (double)N * R, [&] { k_dual<KoalaBear, 64, 1><<<N / 2 / 64, 64>>>(in, out, N, R); });
    timeit("koala dual bs128 minb8", (double)N * R, [&] { k_dual<KoalaBear, 128, 8><<<N / 2 / 128, 128>>>(in, out, N, R); });
    return 0;
}
