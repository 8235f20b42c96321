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
    for (int j = 0; j < 4; j++) sums[j] = fadd<F>(fadd<F>(s[j], s[4 + j]), fadd<F>(s[8 + j], s[12 + j]));
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = fadd<F>(s[i], sums[i & 3]);
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
    for (int r = 0; r < F::ROUNDS_P; r++) {
        s[0] = sbox<F>(fadd<F>(s[0], k.int_rc[r]));
        uint32_t sum = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) sum = fadd<F>(sum, s[i]);
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = fadd<F>(sum, fmul<F>(k.diag[i], s[i]));
    }
#pragma unroll 1
    for (int r = 4; r < 8; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = sbox<F>(fadd<F>(s[i], k.ext_rc[16 * r + i]));
        external_linear<F>(s);
    }
}

#if defined(__CUDACC__)
template <class F>
__device__ __forceinline__ void poseidon2_permute(uint32_t* s) {
    poseidon2_permute_with<F>(s, c_p2[FieldId<F>::value]);
}
#endif

}  // namespace p3r
