// Poseidon2 width-16 permutation throughput (one thread per permutation) as used by k_hash_rows / k_compress.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../plonky3-recursion_b200/csrc/poseidon2.cuh"
using namespace p3r;

template <class F, int BS, int MINB>
__global__ void __launch_bounds__(BS, MINB) k_perm(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, int reps) {
    uint32_t i = blockIdx.x * BS + threadIdx.x;
    if (i >= n) return;
    uint32_t st[16];
#pragma unroll
    for (int k = 0; k < 16; k++) st[k] = in[(size_t)k * n + i];
    for (int r = 0; r < reps; r++) poseidon2_permute<F>(st);
    uint4* o = reinterpret_cast<uint4*>(out + (size_t)i * 8);
    o[0] = make_uint4(st[0], st[1], st[2], st[3]);
    o[1] = make_uint4(st[4], st[5], st[6], st[7]);
}

// Latency of ONE permutation: a single warp runs `reps` dependent permutations (cooperative: 2 states per warp, 16 lanes each;
// serial: 32 states per warp, one per thread). clock64-based, result in cycles per permutation.
template <class F, bool COOP>
__global__ void k_latency(uint32_t* out, const Poseidon2Consts* gk, int reps, long long* cycles) {
    uint32_t lane = threadIdx.x & 31u;
    long long t0, t1;
    if (COOP) {
        P2Lane c = p2_lane_consts<F>(gk, lane & 15u);
        uint32_t x = lane * 7 + 1;
        x = p2_coop_permute<F>(x, lane, c);
        t0 = clock64();
        for (int r = 0; r < reps; r++) x = p2_coop_permute<F>(x, lane, c);
        t1 = clock64();
        out[threadIdx.x] = x;
    } else {
        uint32_t st[16];
#pragma unroll
        for (int k = 0; k < 16; k++) st[k] = lane * 16 + k;
        poseidon2_permute<F>(st);
        t0 = clock64();
        for (int r = 0; r < reps; r++) poseidon2_permute<F>(st);
        t1 = clock64();
        out[threadIdx.x] = st[0] ^ st[5];
    }
    if (threadIdx.x == 0) *cycles = (t1 - t0) / reps;
}

template <class F, int BS, int MINB>
static void run(const char* name, uint32_t* in, uint32_t* out, uint32_t n, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_perm<F, BS, MINB><<<(n + BS - 1) / BS, BS>>>(in, out, n, reps);
    cudaDeviceSynchronize();
    float best = 1e9;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k_perm<F, BS, MINB><<<(n + BS - 1) / BS, BS>>>(in, out, n, reps);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    printf("%-34s n=%8u reps=%d  %8.3f ms  %6.2f perms/ns  %s\n", name, n, reps, best, (double)n * reps / (best * 1e6), err ? cudaGetErrorString(err) : "");
}

int main() {
    Poseidon2Consts h[2];
    for (int f = 0; f < 2; f++) {
        uint32_t x = 12345 + f;
        uint32_t* w = reinterpret_cast<uint32_t*>(&h[f]);
        for (size_t i = 0; i < sizeof(Poseidon2Consts) / 4; i++) { x = x * 1664525u + 1013904223u; w[i] = x % 0x78000001u; }
    }
    cudaMemcpyToSymbol(c_p2, h, sizeof(h));
    const uint32_t N = 1u << 20;
    uint32_t *in, *out;
    cudaMalloc(&in, (size_t)N * 16 * 4);
    cudaMalloc(&out, (size_t)N * 8 * 4);
    cudaMemset(in, 1, (size_t)N * 16 * 4);
    for (uint32_t n : {1u << 20, 1u << 18, 1u << 17, 1u << 16, 1u << 14}) {
        run<KoalaBear, 128, 1>("koala bs128", in, out, n, 1);
        run<KoalaBear, 64, 1>("koala bs64", in, out, n, 1);
        run<KoalaBear, 256, 1>("koala bs256", in, out, n, 1);
        run<KoalaBear, 128, 8>("koala bs128 minb8 (<=64 regs)", in, out, n, 1);
        run<KoalaBear, 128, 12>("koala bs128 minb12 (<=40 regs)", in, out, n, 1);
    }
    {
        Poseidon2Consts* gk;
        cudaMalloc(&gk, sizeof(Poseidon2Consts));
        cudaMemcpy(gk, &h[0], sizeof(Poseidon2Consts), cudaMemcpyHostToDevice);
        long long* cyc;
        cudaMallocManaged(&cyc, 8);
        k_latency<KoalaBear, true><<<1, 32>>>(out, gk, 50, cyc);
        cudaDeviceSynchronize();
        printf("latency koala coop16 (1 warp)   : %lld cycles / permutation\n", *cyc);
        k_latency<KoalaBear, false><<<1, 32>>>(out, gk, 50, cyc);
        cudaDeviceSynchronize();
        printf("latency koala serial (1 warp)   : %lld cycles / permutation\n", *cyc);
        k_latency<BabyBear, true><<<1, 32>>>(out, gk, 50, cyc);
        cudaDeviceSynchronize();
        printf("latency baby  coop16 (1 warp)   : %lld cycles / permutation\n", *cyc);
        k_latency<BabyBear, false><<<1, 32>>>(out, gk, 50, cyc);
        cudaDeviceSynchronize();
        printf("latency baby  serial (1 warp)   : %lld cycles / permutation\n", *cyc);
    }
    run<KoalaBear, 128, 1>("koala bs128 x8 reps", in, out, 1u << 20, 8);
    run<BabyBear, 128, 1>("baby bs128", in, out, 1u << 20, 1);
    run<BabyBear, 128, 1>("baby bs128 x8 reps", in, out, 1u << 20, 8);
    return 0;
}
