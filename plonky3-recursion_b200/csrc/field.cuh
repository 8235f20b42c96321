// field.cuh — Montgomery 31-bit field + degree-4 binomial extension for sm_100a (K1 in SURVEY.md §2.3).
//
// Replaces p3-monty-31's MontyField31 and p3-field's BinomialExtensionField<F,4> (binomial multiplication restated at
// /root/reference circuit-prover/src/air/alu_air.rs:715-733). All values are Montgomery residues x*2^32 mod P kept in
// [0, P). 31-bit modular integer work on the IMAD pipe: no tensor cores.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define P3R_HD __host__ __device__ __forceinline__
#else
#define P3R_HD inline
#endif

namespace p3r {

// Compile-time field parameters. MU = P^{-1} mod 2^32; R = 2^32 mod P; R2 = 2^64 mod P.
struct KoalaBear {
    static constexpr uint32_t P = 0x7f000001u;
    static constexpr uint32_t MU = 0x81000001u;
    static constexpr uint32_t R = 0x01fffffeu;
    static constexpr uint32_t R2 = 0x17f7efe4u;
    static constexpr uint32_t TWO_ADICITY = 24;
    static constexpr int SBOX = 3;
    static constexpr int ROUNDS_P = 20;
};
struct BabyBear {
    static constexpr uint32_t P = 0x78000001u;
    static constexpr uint32_t MU = 0x88000001u;
    static constexpr uint32_t R = 0x0ffffffeu;
    static constexpr uint32_t R2 = 0x45dddde3u;
    static constexpr uint32_t TWO_ADICITY = 27;
    static constexpr int SBOX = 7;
    static constexpr int ROUNDS_P = 13;
};

template <class F>
P3R_HD uint32_t fadd(uint32_t a, uint32_t b) {
    uint32_t s = a + b, s2 = s - F::P;
    return s2 < s ? s2 : s;  // unsigned min(s, s - P)
}
template <class F>
P3R_HD uint32_t fsub(uint32_t a, uint32_t b) {
    uint32_t d = a - b, d2 = d + F::P;
    return d2 < d ? d2 : d;  // a >= b: d < P <= d + P (no wrap) -> d; a < b: d wrapped (>= 2^32 - P), d + P wraps to the answer
}
template <class F>
P3R_HD uint32_t fneg(uint32_t a) {
    return a ? F::P - a : 0u;
}
// Montgomery product in the "positive" form: m = lo(a*b) * (-P^-1), (a*b + m*P) has a zero low word, so the result is the
// high word of ONE multiply-add (IMAD.WIDE with a 64-bit addend) and lies in [0, 2P): 5 SASS instructions
// (IMAD.WIDE, IMAD, IMAD.WIDE, IADD, VIMNMX) instead of 7 for the subtractive form with its 64-bit compare.
// On the device the high word is taken with an explicit carry chain (mad.lo.cc / madc.hi): ptxas then emits exactly
// IMAD.WIDE.U32, IMAD, IMAD.HI.U32 (64-bit addend), VIADDMNMX.U32 — 4 instructions. Written as a 64-bit C expression it
// adds a dead "+ carry-out" IADD3/UMOV pair per product (checked with cuobjdump -sass, CUDA 12.9).
template <class F>
P3R_HD uint32_t fmul(uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a * b;
    uint32_t m = (uint32_t)t * (0u - F::MU);
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("{ .reg .u32 l2;\n\tmad.lo.cc.u32 l2, %1, %2, %3;\n\tmadc.hi.u32 %0, %1, %2, %4; }"
        : "=r"(r)
        : "r"(m), "r"(F::P), "r"((uint32_t)t), "r"((uint32_t)(t >> 32)));
#else
    uint64_t u = (uint64_t)m * F::P + t;  // < 2^63, low 32 bits are zero
    uint32_t r = (uint32_t)(u >> 32);     // < 2P
#endif
    uint32_t r2 = r - F::P;
    return r2 < r ? r2 : r;               // min(r, r - P) as unsigned: r - P wraps above r exactly when r < P
}
// Reduce a 64-bit accumulator of at most FOUR Montgomery products (t <= 4(P-1)^2, so its high word is < 2P) to the
// Montgomery residue of t / 2^32: one unsigned min brings the high word below P, then the same positive-form reduction as
// fmul (IMAD, IMAD.HI with the 64-bit addend, VIADDMNMX) — 5 instructions instead of ~12 for a generic 64-bit reduction.
template <class F>
P3R_HD uint32_t fred64(uint64_t t) {
    uint32_t lo = (uint32_t)t, hi = (uint32_t)(t >> 32);
    uint32_t h2 = hi - F::P;
    hi = h2 < hi ? h2 : hi;  // hi < 2P  ->  [0, P): t' = hi * 2^32 + lo < P * 2^32, t' = t (mod P)
    uint32_t m = lo * (0u - F::MU);
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("{ .reg .u32 l2;\n\tmad.lo.cc.u32 l2, %1, %2, %3;\n\tmadc.hi.u32 %0, %1, %2, %4; }"
        : "=r"(r)
        : "r"(m), "r"(F::P), "r"(lo), "r"(hi));
#else
    uint32_t r = (uint32_t)(((uint64_t)m * F::P + (((uint64_t)hi << 32) | lo)) >> 32);
#endif
    uint32_t r2 = r - F::P;
    return r2 < r ? r2 : r;
}
template <class F>
P3R_HD uint32_t to_monty(uint32_t canonical) {
    return fmul<F>(canonical % F::P, F::R2);
}
template <class F>
P3R_HD uint32_t from_monty(uint32_t m) {
    return fmul<F>(m, 1u);
}
template <class F>
P3R_HD uint32_t fpow(uint32_t a, uint64_t e) {
    uint32_t r = F::R;
    while (e) {
        if (e & 1) r = fmul<F>(r, a);
        a = fmul<F>(a, a);
        e >>= 1;
    }
    return r;
}
template <class F>
P3R_HD uint32_t finv(uint32_t a) {
    return fpow<F>(a, (uint64_t)F::P - 2);
}

// ---- degree-4 binomial extension, x^4 = W (W passed in Montgomery form) ----
struct Ext4 {
    uint32_t c[4];
};
P3R_HD Ext4 ext_zero() { return Ext4{{0, 0, 0, 0}}; }
template <class F>
P3R_HD Ext4 ext_one() {
    return Ext4{{F::R, 0, 0, 0}};
}
template <class F>
P3R_HD Ext4 ext_lift(uint32_t a) {
    return Ext4{{a, 0, 0, 0}};
}
template <class F>
P3R_HD Ext4 eadd(const Ext4& a, const Ext4& b) {
    return Ext4{{fadd<F>(a.c[0], b.c[0]), fadd<F>(a.c[1], b.c[1]), fadd<F>(a.c[2], b.c[2]), fadd<F>(a.c[3], b.c[3])}};
}
template <class F>
P3R_HD Ext4 esub(const Ext4& a, const Ext4& b) {
    return Ext4{{fsub<F>(a.c[0], b.c[0]), fsub<F>(a.c[1], b.c[1]), fsub<F>(a.c[2], b.c[2]), fsub<F>(a.c[3], b.c[3])}};
}
template <class F>
P3R_HD Ext4 eneg(const Ext4& a) {
    return Ext4{{fneg<F>(a.c[0]), fneg<F>(a.c[1]), fneg<F>(a.c[2]), fneg<F>(a.c[3])}};
}
template <class F>
P3R_HD Ext4 emul_base(const Ext4& a, uint32_t b) {
    return Ext4{{fmul<F>(a.c[0], b), fmul<F>(a.c[1], b), fmul<F>(a.c[2], b), fmul<F>(a.c[3], b)}};
}
template <class F>
P3R_HD Ext4 eadd_base(const Ext4& a, uint32_t b) {
    return Ext4{{fadd<F>(a.c[0], b), a.c[1], a.c[2], a.c[3]}};
}
template <class F>
P3R_HD Ext4 esub_base(const Ext4& a, uint32_t b) {
    return Ext4{{fsub<F>(a.c[0], b), a.c[1], a.c[2], a.c[3]}};
}
// Schoolbook product with delayed reduction. The wrap-around operands are multiplied by W first (b_i*W for i = 1..3), so
// every output coefficient is ONE sum of four products (< 4P^2 < 2^64) and one fred64: 16 wide products, 3 products by W and
// 4 reductions (the form with separately reduced wrap-around sums needs 7 reductions and 3 more products).
template <class F>
P3R_HD Ext4 emul(const Ext4& a, const Ext4& b, uint32_t w) {
    const uint32_t bw1 = fmul<F>(b.c[1], w), bw2 = fmul<F>(b.c[2], w), bw3 = fmul<F>(b.c[3], w);
    Ext4 r;
    r.c[0] = fred64<F>((uint64_t)a.c[0] * b.c[0] + (uint64_t)a.c[1] * bw3 + (uint64_t)a.c[2] * bw2 + (uint64_t)a.c[3] * bw1);
    r.c[1] = fred64<F>((uint64_t)a.c[0] * b.c[1] + (uint64_t)a.c[1] * b.c[0] + (uint64_t)a.c[2] * bw3 + (uint64_t)a.c[3] * bw2);
    r.c[2] = fred64<F>((uint64_t)a.c[0] * b.c[2] + (uint64_t)a.c[1] * b.c[1] + (uint64_t)a.c[2] * b.c[0] + (uint64_t)a.c[3] * bw3);
    r.c[3] = fred64<F>((uint64_t)a.c[0] * b.c[3] + (uint64_t)a.c[1] * b.c[2] + (uint64_t)a.c[2] * b.c[1] + (uint64_t)a.c[3] * b.c[0]);
    return r;
}
// Inverse via the tower F -> F(y = x^2) -> F(x): a*(A - Bx) = A^2 - y*B^2 =: n0 + n1*y, 1/(n0+n1 y) = (n0 - n1 y)/(n0^2 - W n1^2).
template <class F>
P3R_HD Ext4 einv(const Ext4& a, uint32_t w) {
    uint32_t a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3];
    uint32_t a1a3 = fmul<F>(a1, a3);
    uint32_t n0 = fsub<F>(fadd<F>(fmul<F>(a0, a0), fmul<F>(w, fmul<F>(a2, a2))), fmul<F>(w, fadd<F>(a1a3, a1a3)));
    uint32_t a0a2 = fmul<F>(a0, a2);
    uint32_t n1 = fsub<F>(fadd<F>(a0a2, a0a2), fadd<F>(fmul<F>(a1, a1), fmul<F>(w, fmul<F>(a3, a3))));
    uint32_t d = fsub<F>(fmul<F>(n0, n0), fmul<F>(w, fmul<F>(n1, n1)));
    uint32_t di = finv<F>(d);
    uint32_t m0 = fmul<F>(n0, di), m1 = fneg<F>(fmul<F>(n1, di));
    Ext4 conj{{a0, fneg<F>(a1), a2, fneg<F>(a3)}};
    Ext4 m{{m0, 0, m1, 0}};
    return emul<F>(conj, m, w);
}
template <class F>
P3R_HD Ext4 epow(Ext4 a, uint64_t e, uint32_t w) {
    Ext4 r = ext_one<F>();
    while (e) {
        if (e & 1) r = emul<F>(r, a, w);
        a = emul<F>(a, a, w);
        e >>= 1;
    }
    return r;
}

P3R_HD uint32_t bitrev32(uint32_t x, uint32_t bits) {
#if defined(__CUDA_ARCH__)
    return bits ? (__brev(x) >> (32 - bits)) : 0u;
#else
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
#endif
}

}  // namespace p3r
