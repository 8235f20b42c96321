// ntt_col.cuh — whole-column NTT kernels for the batched coset LDE (K4), columns of 2^5 .. 2^15 rows.
//
// Replaces p3-dft's Radix2DitParallel::coset_lde_batch ([P3-EXT]; reference call sites circuit-prover/src/config.rs:129-136,
// recursion/examples/common/mod.rs:464-486) for every table of the recursion layer that fits a CTA's shared memory.
// One CTA owns one column (inverse) or one (column, coset) pair (forward); the column is read from HBM/L2 once, all log2(n)
// butterfly stages run on the shared-memory copy, and the result is written once:
//   inverse : decimation in frequency, natural order in -> bit-reversed coefficients out, scaled by 1/n;
//   forward : decimation in time on the bit-reversed coefficients -> evaluations on the coset c*H_n, written to bit-reversed
//             rows of coset block j. The coset shift is folded into the twiddles (stage s uses c^(n/2^(s+1)) * w^e), so there
//             is no per-element scaling pass: p(c*x) = p_even(c^2 x^2) + c*x*p_odd(c^2 x^2).
// Stages are grouped in registers: the first five (strides 1..16) as one radix-32 group on 32 consecutive elements per thread
// with 31 precomputed twiddles held in registers, the rest as radix-8 / radix-16 groups (one table twiddle per group, the
// others derived). The kernel is bound by the integer-multiplier pipe (a Montgomery product is 10 of its cycles), so what
// matters is products per element: 80/32 in the first group, 19/8 or 48/16 in the others.
// Shared-memory layout: word x lives at sigma(x) = x ^ (((x >> 5) & 7) << 2), a 16-byte-granular XOR swizzle that makes the
// thread-contiguous 128-bit accesses of the radix-32 group and the strided accesses of the other groups conflict-free.
#pragma once
#include <cuda_runtime.h>

#include "field.cuh"
#if defined(P3R_NTT_ALU_ADDS)
#include "poseidon2.cuh"   // c_p2[..].zero: a run-time zero in constant memory
#endif

namespace p3r {

// Butterfly addition / subtraction. ptxas balances integer additions between the ALU pipe (IADD3) and the multiplier pipe
// (IMAD.IADD) as if IMAD.WIDE / IMAD.HI cost one slot; they cost two, so in these loops the multiplier pipe is the crowded one
// (scripts/sass_pipe_model.py: radix-32 loop 1213 multiplier-pipe cycles against 584 ALU). With P3R_NTT_ALU_ADDS the additions
// take a third, run-time-zero operand, which only an IADD3 can encode: 1213 -> 967 / 584 -> 818 cycles in that loop, same
// instruction count and registers (static model; to be measured on the GPU before it becomes the default).
#if defined(P3R_NTT_ALU_ADDS)
template <class F>
__device__ __forceinline__ uint32_t ntt_add(uint32_t a, uint32_t b) {
    uint32_t s = a + b + c_p2[FieldId<F>::value].zero, s2 = s - F::P;
    return s2 < s ? s2 : s;
}
template <class F>
__device__ __forceinline__ uint32_t ntt_sub(uint32_t a, uint32_t b) {
    uint32_t d = a - b + c_p2[FieldId<F>::value].zero, d2 = d + F::P;
    return d2 < d ? d2 : d;
}
#else
template <class F>
__device__ __forceinline__ uint32_t ntt_add(uint32_t a, uint32_t b) {
    return fadd<F>(a, b);
}
template <class F>
__device__ __forceinline__ uint32_t ntt_sub(uint32_t a, uint32_t b) {
    return fsub<F>(a, b);
}
#endif

constexpr uint32_t COL_MIN_LOG = 5, COL_MAX_LOG = 15, COL_TOP_MAX_LOG = 27;  // taller columns: k_ntt_top passes above stage 15

struct ColJob {
    const uint32_t* src;
    uint32_t* dst;
    uint64_t src_col_stride, dst_col_stride;  // elements between columns
    uint64_t dst_coset_stride;                // forward: elements between coset blocks
    uint64_t src_coset_stride;                // k_ntt_top forward only
    uint32_t natural_out;                     // forward: leave the result in natural order (stages >= 15 follow in k_ntt_top)
    uint32_t log_n, n_cols, n_cosets;
    uint32_t cols_per_cta;                    // 2^(COL_MAX_LOG - log_n): every CTA works on up to 2^15 elements
    uint32_t cta_begin;                       // first flat CTA of this job; CTA order: (coset, column group)
    uint32_t n_inv;                           // inverse: 1/n (Montgomery)
    uint32_t q[3];                            // plan: stages of the groups after the radix-32 one, bottom-up (0 = none)
    const uint32_t* ctab;                     // per coset COL_CTAB words: 31 first-group twiddles, then C_s for s < 28
    const uint32_t* tws;                      // per-stage compact twiddle tables (w_{2^(s+1)}^e at (2^s - 1) + e)
    uint32_t r4, r8, r8_3;                    // w_4, w_8, w_8^3
    uint32_t r16[8];                          // w_16^e, e < 8
};
constexpr uint32_t COL_CTAB = 31 + 28;

__device__ __forceinline__ uint32_t col_sigma(uint32_t x) { return x ^ (((x >> 5) & 7u) << 2); }

// Butterflies of Q consecutive stages on 2^Q register values; W = twiddle of the group's top stage for this task.
// FWD: DIT  (u, v) -> (u + T v, u - T v), stages ascending;   !FWD: DIF  (u, v) -> (u + v, (u - v) T), stages descending.
// rt[u][e] = w_{2^(u+1)}^(+-e): the fixed roots that combine with powers of W.
template <class F, int Q, bool FWD>
__device__ __forceinline__ void col_group(uint32_t (&v)[1 << Q], uint32_t W, const ColJob& a) {
    // tw[u][e], e < 2^u: twiddle of stage u (relative) for register pairs whose low u bits are e
    uint32_t tw[Q][1 << (Q - 1)];
    {
        uint32_t p = W;  // W^(2^(Q-1-u)) for u = Q-1 down to 0
#pragma unroll
        for (int u = Q - 1; u >= 0; u--) {
            tw[u][0] = p;
            if (u > 0) p = fmul<F>(p, p);
        }
    }
    if (Q >= 2) {
        const uint32_t i4 = FWD ? a.r4 : fneg<F>(a.r4);  // w_4^(+-1)
        tw[1][1] = fmul<F>(tw[1][0], i4);
        if (Q >= 3) {
            const uint32_t i8 = FWD ? a.r8 : fneg<F>(a.r8_3);   // w_8^(+-1)
            const uint32_t i83 = FWD ? a.r8_3 : fneg<F>(a.r8);  // w_8^(+-3)
            tw[2][1] = fmul<F>(tw[2][0], i8);
            tw[2][2] = fmul<F>(tw[2][0], i4);
            tw[2][3] = fmul<F>(tw[2][0], i83);
        }
        if (Q >= 4) {
#pragma unroll
            for (int e = 1; e < 8; e++) {
                const uint32_t rt = FWD ? a.r16[e] : fneg<F>(a.r16[8 - e]);  // w_16^(+-e)
                tw[Q >= 4 ? 3 : 0][e] = fmul<F>(tw[Q >= 4 ? 3 : 0][0], rt);
            }
        }
    }
#pragma unroll
    for (int step = 0; step < Q; step++) {
        const int u = FWD ? step : (Q - 1 - step);
#pragma unroll
        for (int k = 0; k < (1 << Q); k++) {
            if (k & (1 << u)) continue;
            const int k1 = k | (1 << u);
            const uint32_t t = tw[u][k & ((1 << u) - 1)];
            const uint32_t x = v[k], y = v[k1];
            if (FWD) {
                const uint32_t ty = fmul<F>(y, t);
                v[k] = ntt_add<F>(x, ty);
                v[k1] = ntt_sub<F>(x, ty);
            } else {
                v[k] = ntt_add<F>(x, y);
                v[k1] = fmul<F>(ntt_sub<F>(x, y), t);
            }
        }
    }
}

// One group of Q stages at stage offset j >= 5 over the whole column in shared memory.
//   FWD && TO_GLOBAL (top group, j + Q == r): the 2^Q outputs of a task are 2^Q consecutive bit-reversed rows: stored
//   straight to HBM as full 32/64-byte runs.   !FWD && FROM_GLOBAL (top group): inputs are loaded from HBM (coalesced).
template <class F, int Q, bool FWD, bool GLOBAL>
__device__ __forceinline__ void col_stage_group(uint32_t* sm, const ColJob& a, uint32_t log_e, uint32_t j, uint32_t ctop,
                                                const uint32_t* __restrict__ gsrc, uint32_t* __restrict__ gdst, uint32_t n_elems) {
    const uint32_t tasks = n_elems >> Q;
    (void)log_e;
    const uint32_t* tab = a.tws + ((1u << (j + Q - 1)) - 1);
    for (uint32_t tau = threadIdx.x; tau < tasks; tau += blockDim.x) {
        const uint32_t lo = tau & ((1u << j) - 1), hi = tau >> j;
        const uint32_t base = (hi << (j + Q)) | lo;
        uint32_t W;
        if (FWD) W = fmul<F>(__ldg(tab + lo), ctop);
        else W = lo ? fneg<F>(__ldg(tab + ((1u << (j + Q - 1)) - lo))) : F::R;
        uint32_t v[1 << Q];
        const uint32_t sbase = col_sigma(base);
        if (!FWD && GLOBAL) {
#pragma unroll
            for (int k = 0; k < (1 << Q); k++) v[k] = __ldg(gsrc + (size_t)hi * a.src_col_stride + lo + ((uint32_t)k << j));
        } else {
#pragma unroll
            for (int k = 0; k < (1 << Q); k++) v[k] = sm[sbase ^ col_sigma((uint32_t)k << j)];
        }
        col_group<F, Q, FWD>(v, W, a);
        if (FWD && GLOBAL) {
            // natural index i = (k << j) | lo  ->  bit-reversed row (rev_j(lo) << Q) | rev_Q(k)
            uint32_t* o = gdst + (size_t)hi * a.dst_col_stride + ((size_t)(__brev(lo) >> (32 - j)) << Q);
            if (Q >= 2) {
#pragma unroll
                for (int g = 0; g < (1 << Q) / 4; g++) {
                    uint4 w4;
                    w4.x = v[__brev(4 * g + 0) >> (32 - Q)];
                    w4.y = v[__brev(4 * g + 1) >> (32 - Q)];
                    w4.z = v[__brev(4 * g + 2) >> (32 - Q)];
                    w4.w = v[__brev(4 * g + 3) >> (32 - Q)];
                    reinterpret_cast<uint4*>(o)[g] = w4;
                }
            } else {
                reinterpret_cast<uint2*>(o)[0] = make_uint2(v[0], v[1]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < (1 << Q); k++) sm[sbase ^ col_sigma((uint32_t)k << j)] = v[k];
        }
    }
}

template <class F, bool FWD, bool GLOBAL>
__device__ __forceinline__ void col_dispatch_group(uint32_t q, uint32_t* sm, const ColJob& a, uint32_t r, uint32_t j, uint32_t ctop,
                                                   const uint32_t* gsrc, uint32_t* gdst, uint32_t n_elems) {
    if (q == 4) col_stage_group<F, 4, FWD, GLOBAL>(sm, a, r, j, ctop, gsrc, gdst, n_elems);
    else if (q == 3) col_stage_group<F, 3, FWD, GLOBAL>(sm, a, r, j, ctop, gsrc, gdst, n_elems);
    else if (q == 2) col_stage_group<F, 2, FWD, GLOBAL>(sm, a, r, j, ctop, gsrc, gdst, n_elems);
    else col_stage_group<F, 1, FWD, GLOBAL>(sm, a, r, j, ctop, gsrc, gdst, n_elems);
}

// Radix-32 group on the 32 consecutive elements [32 t, 32 t + 32) of every task t (stages 0..4). K = the 31 twiddles,
// K[(1 << s) - 1 + e] for stage s, e < 2^s. Inverse: results are scaled by n_inv.
template <class F, bool FWD>
__device__ __forceinline__ void col_radix32(uint32_t* sm, uint32_t n_elems, const uint32_t (&K)[31], uint32_t n_inv) {
    const uint32_t tasks = n_elems >> 5;
    for (uint32_t t = threadIdx.x; t < tasks; t += blockDim.x) {
        uint32_t v[32];
        uint4* row = reinterpret_cast<uint4*>(sm + 32 * t);
        const uint32_t sw = t & 7u;  // sigma permutes the eight 16-byte chunks of this 128-byte row by xor with (t & 7)
#pragma unroll
        for (int qd = 0; qd < 8; qd++) {
            uint4 w4 = row[qd ^ sw];
            v[4 * qd] = w4.x, v[4 * qd + 1] = w4.y, v[4 * qd + 2] = w4.z, v[4 * qd + 3] = w4.w;
        }
#pragma unroll
        for (int step = 0; step < 5; step++) {
            const int s = FWD ? step : (4 - step);
            const int m = 1 << s;
#pragma unroll
            for (int c = 0; c < 32; c++) {
                if (c & m) continue;
                const uint32_t tw = K[m - 1 + (c & (m - 1))];
                const uint32_t x = v[c], y = v[c + m];
                if (FWD) {
                    const uint32_t ty = fmul<F>(y, tw);
                    v[c] = ntt_add<F>(x, ty);
                    v[c + m] = ntt_sub<F>(x, ty);
                } else {
                    v[c] = ntt_add<F>(x, y);
                    // the inverse twiddle of e == 0 is 1: no product
                    v[c + m] = (c & (m - 1)) ? fmul<F>(ntt_sub<F>(x, y), tw) : ntt_sub<F>(x, y);
                }
            }
        }
        if (!FWD) {
#pragma unroll
            for (int c = 0; c < 32; c++) v[c] = fmul<F>(v[c], n_inv);
        }
#pragma unroll
        for (int qd = 0; qd < 8; qd++) row[qd ^ sw] = make_uint4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
    }
}

// Coalesced copy between HBM (natural layout) and the swizzled shared-memory column, 16 bytes per thread and step.
// Column c of the CTA's group occupies shared-memory words [c << r, (c + 1) << r).
__device__ __forceinline__ void col_copy_in(uint32_t* sm, const uint32_t* __restrict__ g, uint64_t col_stride, uint32_t r,
                                            uint32_t n_elems) {
    uint4* s4 = reinterpret_cast<uint4*>(sm);
    for (uint32_t i = threadIdx.x; i < n_elems / 4; i += blockDim.x) {
        const uint32_t c = i >> (r - 2), x4 = i & ((1u << (r - 2)) - 1);
        s4[i ^ ((i >> 3) & 7u)] = __ldg(reinterpret_cast<const uint4*>(g + (size_t)c * col_stride) + x4);
    }
}
__device__ __forceinline__ void col_copy_out(const uint32_t* sm, uint32_t* __restrict__ g, uint64_t col_stride, uint32_t r,
                                             uint32_t n_elems) {
    const uint4* s4 = reinterpret_cast<const uint4*>(sm);
    for (uint32_t i = threadIdx.x; i < n_elems / 4; i += blockDim.x) {
        const uint32_t c = i >> (r - 2), x4 = i & ((1u << (r - 2)) - 1);
        reinterpret_cast<uint4*>(g + (size_t)c * col_stride)[x4] = s4[i ^ ((i >> 3) & 7u)];
    }
}

template <class F, bool FWD>
__global__ void __launch_bounds__(512) k_ntt_col(const ColJob* __restrict__ jobs, uint32_t n_jobs) {
    extern __shared__ __align__(16) uint32_t sm[];
    __shared__ ColJob a;
    {
        uint32_t j = 0;
        while (j + 1 < n_jobs && blockIdx.x >= jobs[j + 1].cta_begin) j++;
        const uint32_t* srcw = reinterpret_cast<const uint32_t*>(jobs + j);
        uint32_t* dstw = reinterpret_cast<uint32_t*>(&a);
        for (uint32_t i = threadIdx.x; i < sizeof(ColJob) / 4; i += blockDim.x) dstw[i] = srcw[i];
    }
    __syncthreads();
    const uint32_t local = blockIdx.x - a.cta_begin;
    const uint32_t groups = (a.n_cols + a.cols_per_cta - 1) / a.cols_per_cta;
    const uint32_t c0 = (local % groups) * a.cols_per_cta, coset = local / groups;
    const uint32_t r = a.log_n;
    const uint32_t n_elems = min(a.cols_per_cta, a.n_cols - c0) << r;  // columns c0 .. of this CTA, back to back in shared memory
    const uint32_t* src = a.src + (size_t)c0 * a.src_col_stride;
    uint32_t* dst = a.dst + (size_t)c0 * a.dst_col_stride + (FWD ? (size_t)coset * a.dst_coset_stride : 0);
    const uint32_t* ctab = a.ctab + (size_t)(FWD ? coset : 0) * COL_CTAB;
    uint32_t K[31];
#pragma unroll
    for (int i = 0; i < 31; i++) K[i] = __ldg(ctab + i);
    const uint32_t n_groups = (a.q[0] != 0) + (a.q[1] != 0) + (a.q[2] != 0);
    if (FWD) {
        col_copy_in(sm, src, a.src_col_stride, r, n_elems);
        __syncthreads();
        col_radix32<F, true>(sm, n_elems, K, 0);
        __syncthreads();
        uint32_t j = 5;
        for (uint32_t g = 0; g < n_groups; g++) {
            const uint32_t q = a.q[g];
            const uint32_t ctop = __ldg(ctab + 31 + (j + q - 1));
            if (g + 1 == n_groups && !a.natural_out) col_dispatch_group<F, true, true>(q, sm, a, r, j, ctop, nullptr, dst, n_elems);
            else col_dispatch_group<F, true, false>(q, sm, a, r, j, ctop, nullptr, nullptr, n_elems);
            j += q;
            __syncthreads();
        }
        if (a.natural_out) {
            col_copy_out(sm, dst, a.dst_col_stride, r, n_elems);
        } else if (n_groups == 0) {
            // n == 32: bit-reversed copy-out of the radix-32 results
            for (uint32_t i = threadIdx.x; i < n_elems; i += blockDim.x)
                dst[(size_t)(i >> r) * a.dst_col_stride + (__brev(i & 31u) >> 27)] = sm[col_sigma(i)];
        }
    } else {
        uint32_t j = r;
        for (uint32_t g = n_groups; g-- > 0;) {
            const uint32_t q = a.q[g];
            j -= q;
            if (g + 1 == n_groups) col_dispatch_group<F, false, true>(q, sm, a, r, j, 0, src, nullptr, n_elems);
            else col_dispatch_group<F, false, false>(q, sm, a, r, j, 0, nullptr, nullptr, n_elems);
            __syncthreads();
        }
        if (n_groups == 0) {
            col_copy_in(sm, src, a.src_col_stride, r, n_elems);
            __syncthreads();
        }
        col_radix32<F, false>(sm, n_elems, K, a.n_inv);
        __syncthreads();
        col_copy_out(sm, dst, a.dst_col_stride, r, n_elems);
    }
}

// Stages j .. j + Q - 1 (j >= 15) of columns of 2^log_n rows, straight from and to HBM: one task = 2^Q elements 2^j apart,
// consecutive lanes on consecutive rows. Columns taller than 2^15 rows run k_ntt_col on their 2^15-row sub-blocks and one or
// more of these passes for the stages above. Forward: passes ascend, in place in natural order; the pass that ends at
// log_n stores bit-reversed rows into the coset block. Inverse: passes descend from the top (natural in -> natural out),
// k_ntt_col follows.
template <class F, int Q, bool FWD>
__global__ void __launch_bounds__(256) k_ntt_top(ColJob a, uint32_t j, uint32_t bitrev_out) {
    const uint32_t tau = blockIdx.x * blockDim.x + threadIdx.x, col = blockIdx.y, coset = blockIdx.z;
    const uint32_t lo = tau & ((1u << j) - 1), hi = tau >> j;
    const size_t base = ((size_t)hi << (j + Q)) | lo;
    const uint32_t* tab = a.tws + ((1u << (j + Q - 1)) - 1);
    uint32_t W;
    if (FWD) W = fmul<F>(__ldg(tab + lo), __ldg(a.ctab + (size_t)coset * COL_CTAB + 31 + (j + Q - 1)));
    else W = lo ? fneg<F>(__ldg(tab + ((1u << (j + Q - 1)) - lo))) : F::R;
    const uint32_t* src = a.src + (size_t)coset * a.src_coset_stride + (size_t)col * a.src_col_stride + base;
    uint32_t v[1 << Q];
#pragma unroll
    for (int k = 0; k < (1 << Q); k++) v[k] = src[(size_t)k << j];
    col_group<F, Q, FWD>(v, W, a);
    uint32_t* dst = a.dst + (size_t)col * a.dst_col_stride + (size_t)coset * a.dst_coset_stride;
    if (FWD && bitrev_out) {
        // j + Q == log_n, hi == 0: natural index (k << j) | lo -> bit-reversed row (rev_j(lo) << Q) | rev_Q(k)
        uint32_t* o = dst + ((size_t)(__brev(lo) >> (32 - j)) << Q);
        if (Q >= 2) {
#pragma unroll
            for (int g = 0; g < (1 << Q) / 4; g++) {
                uint4 w4;
                w4.x = v[__brev(4 * g + 0) >> (32 - Q)];
                w4.y = v[__brev(4 * g + 1) >> (32 - Q)];
                w4.z = v[__brev(4 * g + 2) >> (32 - Q)];
                w4.w = v[__brev(4 * g + 3) >> (32 - Q)];
                reinterpret_cast<uint4*>(o)[g] = w4;
            }
        } else {
            reinterpret_cast<uint2*>(o)[0] = make_uint2(v[0], v[1]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < (1 << Q); k++) dst[base + ((size_t)k << j)] = v[k];
    }
}

}  // namespace p3r
