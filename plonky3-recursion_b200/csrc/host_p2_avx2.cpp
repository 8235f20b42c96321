// host_p2_avx2.cpp — AVX2 Poseidon2 width-16 permutation for the HOST-side DuplexChallenger (the
// transcript stays on the host; it costs ~490 sequential permutations per layer proof, so its latency is on the proof's
// critical path). Same Montgomery representation and the same results, word for word, as poseidon2_permute_with (the scalar
// twin in poseidon2.cuh) — checked against it at context creation. This translation unit is the only one built with -mavx2;
// p3r.cu calls it only when __builtin_cpu_supports("avx2").
#include <immintrin.h>

#include <cstdint>

#include "poseidon2.cuh"

namespace p3r {

template <class F>
struct Vec {
    static inline __m256i P() { return _mm256_set1_epi32((int)F::P); }
    static inline __m256i add(__m256i a, __m256i b) {
        __m256i t = _mm256_add_epi32(a, b);
        return _mm256_min_epu32(t, _mm256_sub_epi32(t, P()));
    }
    // hi32(l*r) - hi32(q*P) with q = lo32(l*r) * P^-1: a multiple of 2^32 whose high word is the signed Montgomery residue
    static inline __m256i monty_d(__m256i l, __m256i r) {
        __m256i prod = _mm256_mul_epu32(l, r);
        __m256i q = _mm256_mul_epu32(prod, _mm256_set1_epi32((int)F::MU));
        __m256i qp = _mm256_mul_epu32(q, P());
        return _mm256_sub_epi64(prod, qp);
    }
    static inline __m256i mul(__m256i a, __m256i b) {
        __m256i d_evn = monty_d(a, b);
        __m256i d_odd = monty_d(_mm256_srli_epi64(a, 32), _mm256_srli_epi64(b, 32));
        __m256i t = _mm256_blend_epi32(_mm256_srli_epi64(d_evn, 32), d_odd, 0xAA);  // in (-P, P) as signed words
        return _mm256_min_epu32(t, _mm256_add_epi32(t, P()));
    }
    static inline __m256i sbox(__m256i x) {
        __m256i x2 = mul(x, x);
        if (F::SBOX == 3) return mul(x2, x);
        __m256i x3 = mul(x2, x), x4 = mul(x2, x2);
        return mul(x3, x4);
    }
    // M4 on each group of four consecutive words: out_i = (x0+x1+x2+x3) + x_i + 2*x_{i+1}
    static inline __m256i m4(__m256i x) {
        __m256i r1 = _mm256_shuffle_epi32(x, _MM_SHUFFLE(0, 3, 2, 1));
        __m256i r2 = _mm256_shuffle_epi32(x, _MM_SHUFFLE(1, 0, 3, 2));
        __m256i r3 = _mm256_shuffle_epi32(x, _MM_SHUFFLE(2, 1, 0, 3));
        __m256i s = add(x, r1), tot = add(s, add(r2, r3));
        return add(add(tot, s), r1);
    }
    static inline void external(__m256i& lo, __m256i& hi) {
        lo = m4(lo);
        hi = m4(hi);
        __m256i t = add(lo, hi);
        __m256i cs = add(t, _mm256_permute2x128_si256(t, t, 1));  // column sums over the four blocks, in both halves
        lo = add(lo, cs);
        hi = add(hi, cs);
    }
};

template <class F>
static void permute_avx2(uint32_t* st, const Poseidon2Consts& k) {
    using V = Vec<F>;
    __m256i lo = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(st));
    __m256i hi = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(st + 8));
    const __m256i dlo = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(k.diag));
    const __m256i dhi = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(k.diag + 8));
    V::external(lo, hi);
    auto full = [&](int r) {
        lo = V::add(lo, _mm256_loadu_si256(reinterpret_cast<const __m256i*>(k.ext_rc + 16 * r)));
        hi = V::add(hi, _mm256_loadu_si256(reinterpret_cast<const __m256i*>(k.ext_rc + 16 * r + 8)));
        lo = V::sbox(lo);
        hi = V::sbox(hi);
        V::external(lo, hi);
    };
    for (int r = 0; r < 4; r++) full(r);
    // Partial rounds: s[0] lives in a scalar register; the sum of the other 15 words and the 15 diagonal products do not
    // depend on the S-box, so only "S-box -> two scalar additions" is on the round-to-round dependency chain.
    uint32_t s0 = (uint32_t)_mm_cvtsi128_si32(_mm256_castsi256_si128(lo));
    const __m256i zero = _mm256_setzero_si256();
    for (int r = 0; r < F::ROUNDS_P; r++) {
        __m256i t = V::add(_mm256_blend_epi32(lo, zero, 1), hi);
        t = V::add(t, _mm256_shuffle_epi32(t, _MM_SHUFFLE(1, 0, 3, 2)));
        t = V::add(t, _mm256_shuffle_epi32(t, _MM_SHUFFLE(2, 3, 0, 1)));
        t = V::add(t, _mm256_permute2x128_si256(t, t, 1));
        const uint32_t rest = (uint32_t)_mm_cvtsi128_si32(_mm256_castsi256_si128(t));  // s[1] + ... + s[15]
        const __m256i plo = V::mul(lo, dlo), phi = V::mul(hi, dhi);                      // word 0 of plo is unused
        const uint32_t y = sbox<F>(fadd<F>(s0, k.int_rc[r]));
        const uint32_t sum = fadd<F>(rest, y);
        s0 = k.fast_diag ? fsub<F>(rest, y) : fadd<F>(sum, fmul<F>(k.diag[0], y));     // sum + diag[0]*y, diag[0] = -2
        const __m256i sv = _mm256_set1_epi32((int)sum);
        lo = V::add(sv, plo);
        hi = V::add(sv, phi);
    }
    lo = _mm256_blend_epi32(lo, _mm256_set1_epi32((int)s0), 1);
    for (int r = 4; r < 8; r++) full(r);
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(st), lo);
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(st + 8), hi);
}

void host_permute_avx2_koalabear(uint32_t* st, const Poseidon2Consts& k) { permute_avx2<KoalaBear>(st, k); }
void host_permute_avx2_babybear(uint32_t* st, const Poseidon2Consts& k) { permute_avx2<BabyBear>(st, k); }

}  // namespace p3r
