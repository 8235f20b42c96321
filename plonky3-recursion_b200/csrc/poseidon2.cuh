// poseidon2.cuh — Poseidon2 width-16 permutation on sm_100a registers (K5 in SURVEY.md §2.3) plus a host twin used by the
// host-side DuplexChallenger. Replaces p3-poseidon2 / p3-koala-bear / p3-baby-bear `Poseidon2{Koala,Baby}Bear<16>`
// (used by the reference at circuit-prover/src/config.rs:158,181; constants poseidon2-circuit-air/src/public_types.rs:220-226).
// Structure: M_E, 4 x (rc + sbox + M_E), R_P x (rc0 + sbox0 + (1 + diag) layer), 4 x (rc + sbox + M_E), with
// M_E = circ(2*M4, M4, M4, M4), M4 = [[2,3,1,1],[1,2,3,1],[1,1,2,3],[3,1,1,2]].
#pragma once
#include "field.cuh"

namespace p3r {

struct Poseidon2Consts {
    uint32_t ext_rc[8 * 16];   // initial 4 rounds then terminal 4 rounds (Montgomery)
    uint32_t int_rc[32];       // rounds_p entries used
    uint32_t diag[16];
};

#if defined(__CUDACC__)
// One slot per field id (P3R_FIELD_KOALABEAR = 0, P3R_FIELD_BABYBEAR = 1); written by p3r_ctx_create.
__constant__ Poseidon2Consts c_p2[2];
#endif

template <class F>
struct FieldId;
template <>
struct FieldId<KoalaBear> {
    static constexpr int value = 0;
};
template <>
struct FieldId<BabyBear> {
    static constexpr int value = 1;
};

template <class F>
P3R_HD uint32_t sbox(uint32_t x) {
    uint32_t x2 = fmul<F>(x, x);
    if (F::SBOX == 3) return fmul<F>(x2, x);
    uint32_t x3 = fmul<F>(x2, x);
    uint32_t x4 = fmul<F>(x2, x2);
    return fmul<F>(x3, x4);
}

template <class F>
P3R_HD void m4(uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t& x3) {
    uint32_t t01 = fadd<F>(x0, x1);
    uint32_t t23 = fadd<F>(x2, x3);
    uint32_t t0123 = fadd<F>(t01, t23);
    uint32_t t01123 = fadd<F>(t0123, x1);
    uint32_t t01233 = fadd<F>(t0123, x3);
    uint32_t n3 = fadd<F>(t01233, fadd<F>(x0, x0));
    uint32_t n1 = fadd<F>(t01123, fadd<F>(x2, x2));
    uint32_t n0 = fadd<F>(t01123, t01);
    uint32_t n2 = fadd<F>(t01233, t23);
    x0 = n0;
    x1 = n1;
    x2 = n2;
    x3 = n3;
}

template <class F>
P3R_HD void external_linear(uint32_t* s) {
#pragma unroll
    for (int k = 0; k < 4; k++) m4<F>(s[4 * k], s[4 * k + 1], s[4 * k + 2], s[4 * k + 3]);
    uint32_t sums[4];
#pragma unroll
    for (int j = 0; j < 4; j++) sums[j] = fadd<F>(fadd<F>(s[j], s[4 + j]), fadd<F>(s[8 + j], s[12 + j]));  // balanced
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = fadd<F>(s[i], sums[i & 3]);
}

// Internal (partial) round, arranged for instruction-level parallelism: the S-box chain on s[0], the balanced sum of
// s[1..15] and the 15 diagonal products are mutually independent; only two additions follow the S-box.
template <class F>
P3R_HD void internal_round(uint32_t* s, uint32_t rc, const uint32_t* diag) {
    uint32_t a0 = fadd<F>(fadd<F>(s[1], s[2]), fadd<F>(s[3], s[4]));
    uint32_t a1 = fadd<F>(fadd<F>(s[5], s[6]), fadd<F>(s[7], s[8]));
    uint32_t a2 = fadd<F>(fadd<F>(s[9], s[10]), fadd<F>(s[11], s[12]));
    uint32_t a3 = fadd<F>(fadd<F>(s[13], s[14]), s[15]);
    uint32_t rest = fadd<F>(fadd<F>(a0, a1), fadd<F>(a2, a3));
    uint32_t prod[16];
#pragma unroll
    for (int i = 1; i < 16; i++) prod[i] = fmul<F>(diag[i], s[i]);
    uint32_t s0 = sbox<F>(fadd<F>(s[0], rc));
    uint32_t sum = fadd<F>(rest, s0);
    s[0] = fadd<F>(sum, fmul<F>(diag[0], s0));
#pragma unroll
    for (int i = 1; i < 16; i++) s[i] = fadd<F>(sum, prod[i]);
}

template <class F>
P3R_HD void poseidon2_permute_with(uint32_t* s, const Poseidon2Consts& k) {
    external_linear<F>(s);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = sbox<F>(fadd<F>(s[i], k.ext_rc[16 * r + i]));
        external_linear<F>(s);
    }
#pragma unroll 1
    for (int r = 0; r < F::ROUNDS_P; r++) internal_round<F>(s, k.int_rc[r], k.diag);
#pragma unroll 1
    for (int r = 4; r < 8; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = sbox<F>(fadd<F>(s[i], k.ext_rc[16 * r + i]));
        external_linear<F>(s);
    }
}

#if defined(__CUDACC__)
// ---- cooperative permutation: 16 lanes (half a warp) hold one state, lane l owns s[l] -------------------------------
// Used for the small Merkle levels where one-thread-per-permutation is latency-bound (a single warp needs ~10^4 dependent
// instructions per permutation): here a permutation is ~600 warp instructions with the S-boxes of a full round in parallel.
// `rc` = this lane's 8 external round constants, `dg` = this lane's diagonal entry, int_rc from constant memory (uniform).
struct P2Lane {
    uint32_t rc[8];
    uint32_t dg;
};
template <class F>
__device__ __forceinline__ P2Lane p2_lane_consts(const Poseidon2Consts* __restrict__ gk, uint32_t l16) {
    P2Lane c;
#pragma unroll
    for (int r = 0; r < 8; r++) c.rc[r] = __ldg(&gk->ext_rc[16 * r + l16]);
    c.dg = __ldg(&gk->diag[l16]);
    return c;
}
template <class F>
__device__ __forceinline__ uint32_t p2_coop_external(uint32_t x, uint32_t lane) {
    const uint32_t base = lane & ~3u, q = lane & 3u;
    uint32_t a = __shfl_sync(0xffffffffu, x, base | ((q + 1) & 3));
    uint32_t b = __shfl_sync(0xffffffffu, x, base | ((q + 2) & 3));
    uint32_t c = __shfl_sync(0xffffffffu, x, base | ((q + 3) & 3));
    uint32_t t = fadd<F>(x, a), u = fadd<F>(b, c);
    uint32_t o = fadd<F>(fadd<F>(t, t), fadd<F>(a, u));  // 2x + 3a + b + c
    uint32_t y = fadd<F>(o, __shfl_xor_sync(0xffffffffu, o, 4));
    y = fadd<F>(y, __shfl_xor_sync(0xffffffffu, y, 8));
    return fadd<F>(o, y);
}
template <class F>
__device__ __forceinline__ uint32_t p2_coop_permute(uint32_t x, uint32_t lane, const P2Lane& c) {
    const Poseidon2Consts& k = c_p2[FieldId<F>::value];
    const bool lane0 = (lane & 15u) == 0;
    x = p2_coop_external<F>(x, lane);
#pragma unroll
    for (int r = 0; r < 4; r++) x = p2_coop_external<F>(sbox<F>(fadd<F>(x, c.rc[r])), lane);
#pragma unroll 1
    for (int r = 0; r < F::ROUNDS_P; r++) {
        // sum of the 15 untouched lanes proceeds while lane 0's S-box chain runs
        uint32_t rest = lane0 ? 0u : x;
        rest = fadd<F>(rest, __shfl_xor_sync(0xffffffffu, rest, 1));
        rest = fadd<F>(rest, __shfl_xor_sync(0xffffffffu, rest, 2));
        rest = fadd<F>(rest, __shfl_xor_sync(0xffffffffu, rest, 4));
        rest = fadd<F>(rest, __shfl_xor_sync(0xffffffffu, rest, 8));
        uint32_t sb = sbox<F>(fadd<F>(x, k.int_rc[r]));
        x = lane0 ? sb : x;
        uint32_t s0 = __shfl_sync(0xffffffffu, x, lane & ~15u);
        uint32_t sum = fadd<F>(rest, s0);
        x = fadd<F>(sum, fmul<F>(c.dg, x));
    }
#pragma unroll
    for (int r = 4; r < 8; r++) x = p2_coop_external<F>(sbox<F>(fadd<F>(x, c.rc[r])), lane);
    return x;
}
#endif

#if defined(__CUDACC__)
template <class F>
__device__ __forceinline__ void poseidon2_permute(uint32_t* s) {
    poseidon2_permute_with<F>(s, c_p2[FieldId<F>::value]);
}
#endif

}  // namespace p3r
