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
    uint32_t zero;               // always 0, opaque to the compiler: see fadd_alu
    uint32_t fast_diag;          // diag equals the field's structured p3 diagonal (DiagSpec below): products become shifts/adds
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

// The internal-layer diagonals p3 uses for width 16 are +-small integers and +-2^-k, chosen so the products
// need no general multiplication. Entry i: sign, and either an integer multiple m (k == 0) or the power 2^-k.
//   KoalaBear: [-2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, 1/2^8, 1/8, 1/2^24, -1/2^8, -1/8, -1/16, -1/2^24]
//   BabyBear : [-2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, 1/2^8, 1/4, 1/8, 1/2^27, -1/2^8, -1/16, -1/2^27]
// The integer multiplier pipe (IMAD, "fmaheavy") bounds the permutation (ncu: 79 % busy, ALU 50 %), and a Montgomery product
// costs 10 of its cycles, so these 16 products per partial round are the largest avoidable cost.
struct DiagEntry {
    int sign, m, k;
};
template <class F>
struct DiagSpec;
template <>
struct DiagSpec<KoalaBear> {
    static constexpr DiagEntry e[16] = {{-1, 2, 0}, {1, 1, 0},  {1, 2, 0},  {1, 1, 1},   {1, 3, 0},  {1, 4, 0},  {-1, 1, 1}, {-1, 3, 0},
                                        {-1, 4, 0}, {1, 1, 8},  {1, 1, 3},  {1, 1, 24},  {-1, 1, 8}, {-1, 1, 3}, {-1, 1, 4}, {-1, 1, 24}};
};
template <>
struct DiagSpec<BabyBear> {
    static constexpr DiagEntry e[16] = {{-1, 2, 0}, {1, 1, 0},  {1, 2, 0},  {1, 1, 1},   {1, 3, 0},  {1, 4, 0},  {-1, 1, 1}, {-1, 3, 0},
                                        {-1, 4, 0}, {1, 1, 8},  {1, 1, 2},  {1, 1, 3},   {1, 1, 27}, {-1, 1, 8}, {-1, 1, 4}, {-1, 1, 27}};
};
// Canonical value of entry i (host: checked against the caller's diagonal in p3r_ctx_create).
template <class F>
inline uint32_t diag_spec_canonical(int i) {
    const DiagEntry d = DiagSpec<F>::e[i];
    uint64_t v = (uint64_t)d.m % F::P;
    for (int j = 0; j < d.k; j++) v = (v & 1) ? (v + F::P) >> 1 : v >> 1;
    return d.sign > 0 ? (uint32_t)v : (uint32_t)((F::P - v) % F::P);
}
// x * 2^-K mod P for x in [0, P), 1 <= K <= two-adicity: with x = h*2^K + l and 2^-K = -(P-1)/2^K (mod P),
// x * 2^-K = h - l*(P-1)/2^K, which lies in (-P, 2^(31-K)); one unsigned min brings it to [0, P). Montgomery form is
// preserved (the factor is a plain field element).
template <class F, int K>
P3R_HD uint32_t fdiv2k(uint32_t x) {
    if (K == 1) return (x >> 1) + (x & 1u) * ((F::P + 1) >> 1);
    constexpr uint32_t C = (F::P - 1) >> K;
    uint32_t d = (x >> K) - (x & ((1u << K) - 1)) * C;
    uint32_t d2 = d + F::P;
    return d2 < d ? d2 : d;
}
// sum + diag_I * x with the structured diagonal.
template <class F, int I>
P3R_HD uint32_t diag_apply(uint32_t sum, uint32_t x) {
    constexpr DiagEntry d = DiagSpec<F>::e[I];
    uint32_t y = x;
    if (d.k > 0) y = fdiv2k<F, (d.k > 0 ? d.k : 1)>(x);
    else if (d.m == 2) y = fadd<F>(x, x);
    else if (d.m == 3) y = fadd<F>(fadd<F>(x, x), x);
    else if (d.m == 4) {
        y = fadd<F>(x, x);
        y = fadd<F>(y, y);
    }
    return d.sign > 0 ? fadd<F>(sum, y) : fsub<F>(sum, y);
}

// Montgomery product without the final conditional subtraction: result in [0, 2P), valid as ONE operand of a later product.
template <class F>
P3R_HD uint32_t fmul_lazy(uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a * b;
    uint32_t m = (uint32_t)t * (0u - F::MU);
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("{ .reg .u32 l2;\n\tmad.lo.cc.u32 l2, %1, %2, %3;\n\tmadc.hi.u32 %0, %1, %2, %4; }"
        : "=r"(r)
        : "r"(m), "r"(F::P), "r"((uint32_t)t), "r"((uint32_t)(t >> 32)));
    return r;
#else
    return (uint32_t)(((uint64_t)m * F::P + t) >> 32);
#endif
}

template <class F>
P3R_HD uint32_t sbox(uint32_t x) {
    if (F::SBOX == 3) return fmul<F>(fmul_lazy<F>(x, x), x);
    uint32_t x2 = fmul<F>(x, x);
    uint32_t x3 = fmul_lazy<F>(x2, x);
    uint32_t x4 = fmul<F>(x2, x2);
    return fmul<F>(x3, x4);
}

// a + b mod P with the addition forced onto the ALU pipe: `z` is a run-time zero (Poseidon2Consts::zero), and a
// three-input add can only be an IADD3. ptxas otherwise turns about half of the plain additions of the external layer into
// IMAD.IADD on the integer-multiplier pipe, which the Montgomery products already saturate (ncu: fmaheavy 79 %, ALU 50 %).
template <class F>
P3R_HD uint32_t fadd_alu(uint32_t a, uint32_t b, uint32_t z) {
#if defined(P3R_NO_ALU_FORCE)
    (void)z;
    return fadd<F>(a, b);
#else
    uint32_t s = a + b + z, s2 = s - F::P;
    return s2 < s ? s2 : s;
#endif
}
template <class F>
P3R_HD void m4(uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t& x3, uint32_t z) {
    uint32_t t01 = fadd_alu<F>(x0, x1, z);
    uint32_t t23 = fadd_alu<F>(x2, x3, z);
    uint32_t t0123 = fadd_alu<F>(t01, t23, z);
    uint32_t t01123 = fadd_alu<F>(t0123, x1, z);
    uint32_t t01233 = fadd_alu<F>(t0123, x3, z);
    uint32_t n3 = fadd_alu<F>(t01233, fadd<F>(x0, x0), z);
    uint32_t n1 = fadd_alu<F>(t01123, fadd<F>(x2, x2), z);
    uint32_t n0 = fadd_alu<F>(t01123, t01, z);
    uint32_t n2 = fadd_alu<F>(t01233, t23, z);
    x0 = n0;
    x1 = n1;
    x2 = n2;
    x3 = n3;
}

template <class F>
P3R_HD void external_linear(uint32_t* s, uint32_t z = 0) {
#pragma unroll
    for (int k = 0; k < 4; k++) m4<F>(s[4 * k], s[4 * k + 1], s[4 * k + 2], s[4 * k + 3], z);
    uint32_t sums[4];
#pragma unroll
    for (int j = 0; j < 4; j++) sums[j] = fadd<F>(fadd<F>(s[j], s[4 + j]), fadd<F>(s[8 + j], s[12 + j]));  // balanced
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = fadd_alu<F>(s[i], sums[i & 3], z);
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
template <class F, int I>
struct DiagLoop {
    static P3R_HD void run(uint32_t* s, uint32_t sum) {
        s[I] = diag_apply<F, I>(sum, s[I]);
        DiagLoop<F, I + 1>::run(s, sum);
    }
};
template <class F>
struct DiagLoop<F, 16> {
    static P3R_HD void run(uint32_t*, uint32_t) {}
};
// Same round with the structured diagonal: no general product outside the S-box.
template <class F>
P3R_HD void internal_round_fast(uint32_t* s, uint32_t rc) {
    uint32_t a0 = fadd<F>(fadd<F>(s[1], s[2]), fadd<F>(s[3], s[4]));
    uint32_t a1 = fadd<F>(fadd<F>(s[5], s[6]), fadd<F>(s[7], s[8]));
    uint32_t a2 = fadd<F>(fadd<F>(s[9], s[10]), fadd<F>(s[11], s[12]));
    uint32_t a3 = fadd<F>(fadd<F>(s[13], s[14]), s[15]);
    uint32_t rest = fadd<F>(fadd<F>(a0, a1), fadd<F>(a2, a3));
    s[0] = sbox<F>(fadd<F>(s[0], rc));
    uint32_t sum = fadd<F>(rest, s[0]);
    DiagLoop<F, 0>::run(s, sum);
}

template <class F>
P3R_HD void poseidon2_permute_with(uint32_t* s, const Poseidon2Consts& k) {
    external_linear<F>(s, k.zero);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = sbox<F>(fadd<F>(s[i], k.ext_rc[16 * r + i]));
        external_linear<F>(s, k.zero);
    }
    if (k.fast_diag) {
#pragma unroll 1
        for (int r = 0; r < F::ROUNDS_P; r++) internal_round_fast<F>(s, k.int_rc[r]);
    } else {
#pragma unroll 1
        for (int r = 0; r < F::ROUNDS_P; r++) internal_round<F>(s, k.int_rc[r], k.diag);
    }
#pragma unroll 1
    for (int r = 4; r < 8; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = sbox<F>(fadd<F>(s[i], k.ext_rc[16 * r + i]));
        external_linear<F>(s, k.zero);
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
    bool d0_neg2;   // diag[0] == -2 (true for both p3 parameter sets): lane 0's update is rest - s0, no product on the chain
};
template <class F>
__device__ __forceinline__ P2Lane p2_lane_consts(const Poseidon2Consts* __restrict__ gk, uint32_t l16) {
    P2Lane c;
#pragma unroll
    for (int r = 0; r < 8; r++) c.rc[r] = __ldg(&gk->ext_rc[16 * r + l16]);
    c.dg = __ldg(&gk->diag[l16]);
    c.d0_neg2 = __ldg(&gk->diag[0]) == fsub<F>(0u, fadd<F>(F::R, F::R));
    return c;
}
// Sum over the 4 lanes {l, l^m, l^2m, l^3m} in one shuffle round (3 independent shuffles + a depth-2 add tree): two rounds
// reduce 16 lanes, against four dependent shuffle+add steps for the xor butterfly.
template <class F>
__device__ __forceinline__ uint32_t p2_sum4(uint32_t v, uint32_t m) {
    uint32_t a = __shfl_xor_sync(0xffffffffu, v, m);
    uint32_t b = __shfl_xor_sync(0xffffffffu, v, 2 * m);
    uint32_t c = __shfl_xor_sync(0xffffffffu, v, 3 * m);
    return fadd<F>(fadd<F>(v, a), fadd<F>(b, c));
}
template <class F>
__device__ __forceinline__ uint32_t p2_coop_external(uint32_t x, uint32_t lane) {
    const uint32_t base = lane & ~3u, q = lane & 3u;
    uint32_t a = __shfl_sync(0xffffffffu, x, base | ((q + 1) & 3));
    uint32_t b = __shfl_sync(0xffffffffu, x, base | ((q + 2) & 3));
    uint32_t c = __shfl_sync(0xffffffffu, x, base | ((q + 3) & 3));
    uint32_t t = fadd<F>(x, a), u = fadd<F>(b, c);
    uint32_t o = fadd<F>(fadd<F>(t, t), fadd<F>(a, u));  // 2x + 3a + b + c
    return fadd<F>(o, p2_sum4<F>(o, 4));                 // + column sum over the four M4 blocks
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
        // Off the S-box chain: the diagonal products of the 15 untouched lanes and their sum (two shuffle rounds).
        uint32_t prod = fmul<F>(c.dg, x);
        uint32_t rest = p2_sum4<F>(p2_sum4<F>(lane0 ? 0u : x, 1), 4);
        uint32_t sb = sbox<F>(fadd<F>(x, k.int_rc[r]));          // meaningful on lane 0
        uint32_t s0 = __shfl_sync(0xffffffffu, sb, lane & ~15u);
        uint32_t sum = fadd<F>(rest, s0);
        uint32_t x0 = c.d0_neg2 ? fsub<F>(rest, s0) : fadd<F>(sum, fmul<F>(c.dg, s0));  // sum + diag[0]*s0
        x = lane0 ? x0 : fadd<F>(sum, prod);
    }
#pragma unroll
    for (int r = 4; r < 8; r++) x = p2_coop_external<F>(sbox<F>(fadd<F>(x, c.rc[r])), lane);
    return x;
}
#endif

// ---- generic width (16 or 24): the leaf hasher of the `PaddingFreeSponge<Perm24, 24, 16, 8>` configurations
// (circuit/src/ops/poseidon2_perm/config.rs:77-86,124-133: BABY_BEAR_D4_W24 / KOALA_BEAR_D4_W24; the reference's `Config` is generic
// over the hash permutation's width and rate, circuit-prover/src/config.rs:59-74). Constants come from global memory (per
// context, no __constant__ slot), the diagonal is applied with general products: this path serves the width-24 leaf hashing,
// the width-16 hot path keeps its specialised code above.
struct Poseidon2ConstsW {
    uint32_t width, rounds_p;
    uint32_t ext_rc[8 * 24];
    uint32_t int_rc[32];
    uint32_t diag[24];
};
template <class F, int W>
P3R_HD void external_linear_w(uint32_t* s) {
#pragma unroll
    for (int k = 0; k < W / 4; k++) m4<F>(s[4 * k], s[4 * k + 1], s[4 * k + 2], s[4 * k + 3], 0u);
    uint32_t sums[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t t = s[j];
#pragma unroll
        for (int k = 1; k < W / 4; k++) t = fadd<F>(t, s[4 * k + j]);
        sums[j] = t;
    }
#pragma unroll
    for (int i = 0; i < W; i++) s[i] = fadd<F>(s[i], sums[i & 3]);
}
template <class F, int W>
P3R_HD void poseidon2_permute_w(uint32_t* s, const Poseidon2ConstsW* __restrict__ k) {
    external_linear_w<F, W>(s);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < W; i++) s[i] = sbox<F>(fadd<F>(s[i], k->ext_rc[W * r + i]));
        external_linear_w<F, W>(s);
    }
#pragma unroll 1
    for (uint32_t r = 0; r < k->rounds_p; r++) {
        s[0] = sbox<F>(fadd<F>(s[0], k->int_rc[r]));
        uint32_t sum = s[0];
#pragma unroll
        for (int i = 1; i < W; i++) sum = fadd<F>(sum, s[i]);
#pragma unroll
        for (int i = 0; i < W; i++) s[i] = fadd<F>(sum, fmul<F>(k->diag[i], s[i]));
    }
#pragma unroll 1
    for (int r = 4; r < 8; r++) {
#pragma unroll
        for (int i = 0; i < W; i++) s[i] = sbox<F>(fadd<F>(s[i], k->ext_rc[W * r + i]));
        external_linear_w<F, W>(s);
    }
}

#if defined(__CUDACC__)
template <class F>
__device__ __forceinline__ void poseidon2_permute(uint32_t* s) {
    poseidon2_permute_with<F>(s, c_p2[FieldId<F>::value]);
}
#endif

}  // namespace p3r
