// kernels.cuh — all device kernels of the B200 batch-STARK prover (sm_100a, hand-written; INT32/IMAD pipe, no tensor cores).
//
// Device data layout (DESIGN.md "Data layout in HBM"): every committed matrix lives COLUMN-MAJOR in HBM (column c of a
// matrix of height H is the contiguous run d[c*H .. c*H+H)), LDE rows in bit-reversed order (p3's storage convention,
// /root/reference recursion/src/pcs/fri/verifier.rs:921-976). With one thread per row every kernel below reads 128-byte
// coalesced column segments. Extension-field vectors of the FRI phase are AoS (one uint4 per element).
#pragma once
#include <cuda_runtime.h>

#include "../../include/p3r.h"
#include "field.cuh"
#include "poseidon2.cuh"

namespace p3r {

// omega_T^e from the half table tw[k] = omega_T^k (k < T/2), T = 2^logT.
template <class F>
__device__ __forceinline__ uint32_t root_pow(const uint32_t* __restrict__ tw, uint32_t logT, uint64_t e) {
    uint32_t T = 1u << logT, x = (uint32_t)e & (T - 1), half = T >> 1;
    return x >= half ? fneg<F>(__ldg(tw + (x - half))) : __ldg(tw + x);
}

// ------------------------------------------------------------------------------------------------
// Setup kernels
// ------------------------------------------------------------------------------------------------
// tw[k] = w^k for k < n, w given (Montgomery); two-level so each thread does O(log) work.
template <class F>
__global__ void k_powers(uint32_t* out, uint32_t n, uint32_t w, uint32_t first) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = fmul<F>(first, fpow<F>(w, i));
}

// Row-major (h x w) -> column-major (w x h) through a 32x33 shared tile (coalesced on both sides).
static __global__ void k_transpose_in(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t h, uint32_t w) {
    __shared__ uint32_t tile[32][33];
    uint32_t c0 = blockIdx.y * 32, r0 = blockIdx.x * 32;  // rows on grid.x (no 65535 limit)
    for (uint32_t dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        uint32_t r = r0 + dy, c = c0 + threadIdx.x;
        if (r < h && c < w) tile[dy][threadIdx.x] = in[(size_t)r * w + c];
    }
    __syncthreads();
    for (uint32_t dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        uint32_t c = c0 + dy, r = r0 + threadIdx.x;
        if (r < h && c < w) out[(size_t)c * h + r] = tile[threadIdx.x][dy];
    }
}
// Column-major (w x h) -> row-major (h x w); used only by the isolated p3r_coset_lde entry point.
static __global__ void k_transpose_out(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t h, uint32_t w) {
    __shared__ uint32_t tile[32][33];
    uint32_t c0 = blockIdx.y * 32, r0 = blockIdx.x * 32;  // rows on grid.x (no 65535 limit)
    for (uint32_t dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        uint32_t c = c0 + dy, r = r0 + threadIdx.x;
        if (r < h && c < w) tile[dy][threadIdx.x] = in[(size_t)c * h + r];
    }
    __syncthreads();
    for (uint32_t dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        uint32_t r = r0 + dy, c = c0 + threadIdx.x;
        if (r < h && c < w) out[(size_t)r * w + c] = tile[threadIdx.x][dy];
    }
}

// rows [row0, row0 + n_rows) of a column-major device matrix (height h, width w) <- row-major `rows` (p3r_traces_write_rows).
static __global__ void k_scatter_rows(const uint32_t* __restrict__ rows, uint32_t* __restrict__ out, uint32_t h, uint32_t w,
                                      uint32_t row0, uint32_t n_rows) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * w) return;
    const uint32_t r = i / w, c = i % w;
    out[(size_t)c * h + row0 + r] = rows[i];
}

// ------------------------------------------------------------------------------------------------
// K4 (second implementation): multi-pass tile kernel for the batched coset LDE. The product path uses the whole-column
// kernels of ntt_col.cuh for every column of 2^5 rows or more; this kernel serves columns shorter than 32 rows and is the
// independent implementation the parity tests compare against (p3r_set_specialization bit 1).
// Radix-2 NTT passes over shared-memory tiles, butterflies done in REGISTERS in groups of up to three stages (radix-8):
// one shared-memory round trip and one barrier per group instead of per stage.
// A pass executes butterfly stages [s0, s0+r) of a size-2^log_n transform on every column (grid.y) and, for forward passes,
// every output coset (grid.z). Stage s pairs i and i + 2^s with twiddle omega_{2^{s+1}}^{i mod 2^s}.
//   inverse  : decimation-in-frequency, stages descending, natural in -> bit-reversed out, inverse twiddles (no 1/n here);
//   forward  : decimation-in-time, stages ascending, bit-reversed in -> natural out; the first pass multiplies coefficient k
//              by (shift_j^k / n), the last pass stores to bit-reversed positions inside coset block j.
// Tile = 2^r butterfly rows (stride 2^s0) x 2^log_cw consecutive elements; s0 == 0 tiles are contiguous runs.
// Twiddles: `tws` holds, for every stage s, the compact table omega_{2^{s+1}}^e (e < 2^s) at offset 2^s - 1, so the one
// twiddle a thread loads per group is a coalesced read; the other twiddles of the group are W^2, W^4 and W times 4th/8th roots.
// ------------------------------------------------------------------------------------------------
struct NttPass {
    const uint32_t* src;
    uint32_t* dst;
    uint64_t src_col_stride, dst_col_stride;  // elements between columns
    uint64_t dst_coset_stride;                // elements between coset blocks in dst (forward)
    uint32_t log_n, s0, r, log_cw;
    uint32_t forward;        // 0 = inverse DIF, 1 = forward DIT
    uint32_t first, last;    // first / last pass of the transform
    uint32_t log_blowup;     // forward: number of coset blocks = 2^log_blowup
    uint32_t use_g;          // scale by GENERATOR^k (trace-like input domain H_n) or not (input domain already shifted by GENERATOR)
    uint32_t rot;            // extra root-of-unity rotation exponent (mod N) applied per coefficient index
    uint32_t n_inv;          // 1/n (Montgomery), used when use_g == 0
    const uint32_t* tws;     // per-stage compact twiddle tables
    const uint32_t* tw;      // half table of omega_T (for the coset scaling)
    uint32_t logT;
    uint32_t r4, r8, r8_3;   // omega_4, omega_8, omega_8^3 (Montgomery)
    const uint32_t* g_lo;    // g^k / n = g_lo[k & 1023] * g_hi[k >> 10]
    const uint32_t* g_hi;
    uint32_t n_tiles, n_cols, n_cosets;  // CTA decomposition of this job inside a multi-job launch
    uint32_t cta_begin;                  // first flat CTA index of this job
};

// Butterflies of Q consecutive stages on the 2^Q register values v[k] (k = bits j..j+Q-1 of the tile row index).
// W = twiddle of the group's top stage for this thread's (t_lo, column); q4/q8/q83 = 4th/8th roots (already inverted for DIF).
template <class F, int Q, bool FWD>
__device__ __forceinline__ void ntt_group(uint32_t (&v)[8], uint32_t W, uint32_t q4, uint32_t q8, uint32_t q83) {
    uint32_t tw[3][4];
    if (Q == 3) {
        uint32_t W2 = fmul<F>(W, W);
        tw[0][0] = fmul<F>(W2, W2);
        tw[1][0] = W2;
        tw[1][1] = fmul<F>(W2, q4);
        tw[2][0] = W;
        tw[2][1] = fmul<F>(W, q8);
        tw[2][2] = fmul<F>(W, q4);
        tw[2][3] = fmul<F>(W, q83);
    } else if (Q == 2) {
        tw[0][0] = fmul<F>(W, W);
        tw[1][0] = W;
        tw[1][1] = fmul<F>(W, q4);
    } else {
        tw[0][0] = W;
    }
#pragma unroll
    for (int step = 0; step < Q; step++) {
        const int u = FWD ? step : (Q - 1 - step);
#pragma unroll
        for (int k = 0; k < (1 << Q); k++) {
            if (k & (1 << u)) continue;
            const int k1 = k | (1 << u);
            const uint32_t w = tw[u][k & ((1 << u) - 1)];
            uint32_t x = v[k], y = v[k1];
            if (FWD) {
                y = fmul<F>(y, w);
                v[k] = fadd<F>(x, y);
                v[k1] = fsub<F>(x, y);
            } else {
                v[k] = fadd<F>(x, y);
                v[k1] = fmul<F>(fsub<F>(x, y), w);
            }
        }
    }
}

template <class F, int Q, bool FWD>
__device__ __forceinline__ void ntt_group_pass(uint32_t* sm, const NttPass& a, uint32_t j, uint32_t m0, bool contig, uint32_t CWP) {
    const uint32_t R = 1u << a.r, S = 1u << a.s0;
    const uint32_t total = R << a.log_cw;
    const uint32_t tasks = total >> Q;
    const uint32_t s_top = a.s0 + j + Q - 1;
    const uint32_t* tab = a.tws + ((1u << s_top) - 1);
    const uint32_t q4 = FWD ? a.r4 : fneg<F>(a.r4);
    const uint32_t q8 = FWD ? a.r8 : fneg<F>(a.r8_3);
    const uint32_t q83 = FWD ? a.r8_3 : fneg<F>(a.r8);
    for (uint32_t tau = threadIdx.x; tau < tasks; tau += blockDim.x) {
        uint32_t mm, t_lo, t_hi;
        if (contig) {
            t_lo = tau & ((1u << j) - 1);
            uint32_t rest = tau >> j;
            uint32_t hi_cnt_log = a.r - j - Q;
            t_hi = rest & ((1u << hi_cnt_log) - 1);
            mm = rest >> hi_cnt_log;
        } else {
            mm = tau & ((1u << a.log_cw) - 1);
            uint32_t rest = tau >> a.log_cw;
            t_lo = rest & ((1u << j) - 1);
            t_hi = rest >> j;
        }
        const uint32_t tbase = (t_hi << (j + Q)) | t_lo;
        const uint32_t lo_g = contig ? 0u : ((m0 + mm) & (S - 1));
        const uint32_t e = t_lo * S + lo_g;
        uint32_t W;
        if (FWD) W = __ldg(tab + e);
        else W = e ? fneg<F>(__ldg(tab + ((1u << s_top) - e))) : F::R;
        uint32_t v[8];
        uint32_t idx[8];
#pragma unroll
        for (int k = 0; k < (1 << Q); k++) {
            uint32_t t = tbase | ((uint32_t)k << j);
            uint32_t x = contig ? (mm << a.r) + t : t * CWP + mm;
            if (contig) x += x >> 5;
            idx[k] = x;
            v[k] = sm[x];
        }
        ntt_group<F, Q, FWD>(v, W, q4, q8, q83);
#pragma unroll
        for (int k = 0; k < (1 << Q); k++) sm[idx[k]] = v[k];
    }
}

template <class F, bool FWD>
__device__ __forceinline__ void ntt_tile_stages(uint32_t* sm, const NttPass& a, uint32_t m0, bool contig, uint32_t CWP) {
    // forward: groups ascend from stage offset 0; inverse: groups descend from the top
    uint32_t done = 0;
    while (done < a.r) {
        uint32_t q = a.r - done >= 3 ? 3 : a.r - done;
        uint32_t j = FWD ? done : (a.r - done - q);
        if (q == 3) ntt_group_pass<F, 3, FWD>(sm, a, j, m0, contig, CWP);
        else if (q == 2) ntt_group_pass<F, 2, FWD>(sm, a, j, m0, contig, CWP);
        else ntt_group_pass<F, 1, FWD>(sm, a, j, m0, contig, CWP);
        done += q;
        __syncthreads();
    }
}

// One launch runs the same pass level of MANY LDE jobs (all tables of a commit round): flat grid, each CTA finds its job
// through cta_begin. Keeps the number of dependent launches per commit at (inverse passes + forward passes).
template <class F>
__global__ void __launch_bounds__(512) k_ntt_pass(const NttPass* __restrict__ jobs, uint32_t n_jobs) {
    extern __shared__ uint32_t sm[];
    __shared__ NttPass a;
    {
        uint32_t j = 0;
        while (j + 1 < n_jobs && blockIdx.x >= jobs[j + 1].cta_begin) j++;
        const uint32_t* srcw = reinterpret_cast<const uint32_t*>(jobs + j);
        uint32_t* dstw = reinterpret_cast<uint32_t*>(&a);
        for (uint32_t i = threadIdx.x; i < sizeof(NttPass) / 4; i += blockDim.x) dstw[i] = srcw[i];
    }
    __syncthreads();
    const uint32_t local = blockIdx.x - a.cta_begin;
    const uint32_t tile_id = local % a.n_tiles, col = (local / a.n_tiles) % a.n_cols, coset = local / (a.n_tiles * a.n_cols);
    const uint32_t R = 1u << a.r, CW = 1u << a.log_cw, S = 1u << a.s0;
    const bool contig = (a.s0 == 0);          // tile = CW contiguous runs of R elements
    const uint32_t CWP = CW + 1;              // strided layout: padded row stride
    const uint32_t m0 = tile_id * CW;
    const uint32_t* src = a.src + (size_t)col * a.src_col_stride;
    if (a.forward && !a.first) src += (size_t)coset * a.dst_coset_stride;  // in-place chain lives in the coset block
    uint32_t* dst = a.dst + (size_t)col * a.dst_col_stride + (size_t)coset * (a.forward ? a.dst_coset_stride : 0);
    const uint32_t total = R * CW;

    // ---- load ----
    const uint32_t rj = bitrev32(coset, a.log_blowup);
    const uint32_t logN = a.log_n + a.log_blowup;
    for (uint32_t idx = threadIdx.x; idx < total; idx += blockDim.x) {
        uint32_t gi, x;
        if (contig) {
            gi = m0 * R + idx;  // (m0+mm)*R + t with idx = mm*R + t
            x = idx + (idx >> 5);
        } else {
            uint32_t mm = idx & (CW - 1), t = idx >> a.log_cw;
            uint32_t m = m0 + mm, lo = m & (S - 1), hi = m >> a.s0;
            gi = ((hi << a.r) + t) * S + lo;
            x = t * CWP + mm;
        }
        uint32_t v = src[gi];
        if (a.forward && a.first) {
            uint32_t k = bitrev32(gi, a.log_n);  // coefficient index held at position gi
            uint32_t sc = a.use_g ? fmul<F>(__ldg(a.g_lo + (k & 1023)), __ldg(a.g_hi + (k >> 10))) : a.n_inv;
            uint64_t e = ((uint64_t)(rj + a.rot) * k) & ((1ull << logN) - 1);
            sc = fmul<F>(sc, root_pow<F>(a.tw, a.logT, e << (a.logT - logN)));
            v = fmul<F>(v, sc);
        }
        sm[x] = v;
    }
    __syncthreads();

    if (a.forward) ntt_tile_stages<F, true>(sm, a, m0, contig, CWP);
    else ntt_tile_stages<F, false>(sm, a, m0, contig, CWP);

    // ---- store ----
    if (a.forward && a.last) {
        // natural index i = t*S + lo (hi == 0 since s0 + r == log_n)  ->  bitrev(i) = bitrev_s0(lo)*R + bitrev_r(t)
        for (uint32_t idx = threadIdx.x; idx < total; idx += blockDim.x) {
            uint32_t tt = idx & (R - 1), mm = idx >> a.r;
            uint32_t tb = bitrev32(tt, a.r);
            uint32_t x;
            if (contig) {
                x = (mm << a.r) + tb;
                x += x >> 5;
            } else {
                x = tb * CWP + mm;
            }
            uint32_t pos = bitrev32(m0 + mm, a.s0) * R + tt;
            dst[pos] = sm[x];
        }
    } else {
        for (uint32_t idx = threadIdx.x; idx < total; idx += blockDim.x) {
            uint32_t gi, x;
            if (contig) {
                gi = m0 * R + idx;
                x = idx + (idx >> 5);
            } else {
                uint32_t mm = idx & (CW - 1), t = idx >> a.log_cw;
                uint32_t m = m0 + mm, lo = m & (S - 1), hi = m >> a.s0;
                gi = ((hi << a.r) + t) * S + lo;
                x = t * CWP + mm;
            }
            dst[gi] = sm[x];
        }
    }
}

// tws[(2^s - 1) + e] = omega_{2^{s+1}}^e for s < logT, e < 2^s  (w_T = primitive 2^logT-th root, Montgomery).
template <class F>
__global__ void k_stage_twiddles(uint32_t* tws, uint32_t logT, uint32_t w_T) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // linear index into tws, < 2^logT - 1
    if (i >= (1u << logT) - 1) return;
    uint32_t s = 31 - __clz(i + 1);
    uint32_t e = i + 1 - (1u << s);
    tws[i] = fpow<F>(w_T, (uint64_t)e << (logT - s - 1));
}

// ------------------------------------------------------------------------------------------------
// K5: Poseidon2 sponge over rows + 2-to-1 compression tree (MerkleTreeMmcs).
// ------------------------------------------------------------------------------------------------
// Row digests of the matrices of EVERY height of one commit in a single launch (flat grid, each CTA finds its job through
// cta_begin; jobs sorted by decreasing sponge length so the long rows start first). The tallest height's digests are level 0
// of the tree, the others are the digests injected at their level. Hashing the injected rows here instead of inside the
// level's compression kernel gives the whole commit's sponge work to one grid (1.5 waves for the recursion layer) instead of
// a 2^16-thread launch running 23 dependent permutations per thread at 18 % occupancy.
// colptr[i] = pointer to column i (column-major); one thread per row.
struct HashJob {
    const uint32_t* const* colptr;
    uint32_t ncols, n_rows;
    uint32_t* out;          // n_rows x 8 words
    uint32_t cta_begin;
};
template <class F>
__global__ void __launch_bounds__(128) k_hash_rows(const HashJob* __restrict__ jobs, uint32_t n_jobs) {
    uint32_t j = 0;
    while (j + 1 < n_jobs && blockIdx.x >= jobs[j + 1].cta_begin) j++;
    const HashJob job = jobs[j];
    uint32_t r = (blockIdx.x - job.cta_begin) * blockDim.x + threadIdx.x;
    if (r >= job.n_rows) return;
    uint32_t st[16];
#pragma unroll
    for (int i = 0; i < 16; i++) st[i] = 0;
    for (uint32_t c0 = 0; c0 < job.ncols; c0 += 8) {
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (c0 + k < job.ncols) st[k] = __ldg(job.colptr[c0 + k] + r);
        poseidon2_permute<F>(st);
    }
    uint4* o = reinterpret_cast<uint4*>(job.out + (size_t)r * 8);
    o[0] = make_uint4(st[0], st[1], st[2], st[3]);
    o[1] = make_uint4(st[4], st[5], st[6], st[7]);
}
// Same result through a work queue: a work item is 32 consecutive rows of one job (one warp), items are numbered with the
// longest sponges first and every warp of a machine-filling grid takes the next item from an atomic counter when it finishes
// its current one (longest-processing-time-first list scheduling). Kept as the second schedule for the parity tests
// (p3r_set_specialization bit 3): on B200 it is SLOWER than k_hash_rows (hash class 1.44 ms against 1.18 ms per layer proof):
// ptxas gives the persistent loop 40 registers (48 resident warps instead of 64) and puts 48 instead of 28 of a full round's
// additions on the multiplier pipe (478 against 436 multiplier-pipe cycles per round and warp, scripts/sass_pipe_model.py).
// The load imbalance it was written for is handled by the CTA size of k_hash_rows instead (64 rows, p3r.cu commit_tree).
struct HashQueue {
    uint32_t n_jobs, n_items;
    uint32_t counter;        // zero when the launch starts (uploaded with the descriptor)
    uint32_t pad;
};
template <class F>
__global__ void __launch_bounds__(128) k_hash_rows_queue(const HashJob* __restrict__ jobs, HashQueue* __restrict__ q) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_jobs = q->n_jobs, n_items = q->n_items;
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(&q->counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) return;
        uint32_t j = 0;
        while (j + 1 < n_jobs && item >= __ldg(&jobs[j + 1].cta_begin)) j++;   // cta_begin = first item of the job here
        const HashJob job = jobs[j];
        const uint32_t r = (item - job.cta_begin) * 32u + lane;
        if (r >= job.n_rows) continue;
        uint32_t st[16];
#pragma unroll
        for (int i = 0; i < 16; i++) st[i] = 0;
        for (uint32_t c0 = 0; c0 < job.ncols; c0 += 8) {
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (c0 + k < job.ncols) st[k] = __ldg(job.colptr[c0 + k] + r);
            poseidon2_permute<F>(st);
        }
        uint4* o = reinterpret_cast<uint4*>(job.out + (size_t)r * 8);
        o[0] = make_uint4(st[0], st[1], st[2], st[3]);
        o[1] = make_uint4(st[4], st[5], st[6], st[7]);
    }
}
// Rows of a row-major matrix of `w` words (ExtensionMmcs rows of the FRI commit phase, recursion/src/pcs/mmcs.rs:434-441).
template <class F>
__global__ void __launch_bounds__(128) k_hash_rows_rowmajor(const uint32_t* __restrict__ data, uint32_t w, uint32_t n_rows,
                                                             uint32_t* __restrict__ out) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    uint32_t st[16];
#pragma unroll
    for (int i = 0; i < 16; i++) st[i] = 0;
    const uint32_t* row = data + (size_t)r * w;
    for (uint32_t c0 = 0; c0 < w; c0 += 8) {
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (c0 + k < w) st[k] = row[c0 + k];
        poseidon2_permute<F>(st);
    }
    uint4* o = reinterpret_cast<uint4*>(out + (size_t)r * 8);
    o[0] = make_uint4(st[0], st[1], st[2], st[3]);
    o[1] = make_uint4(st[4], st[5], st[6], st[7]);
}
// ---- width-24 leaf hasher: PaddingFreeSponge<Perm24, 24, 16, 8> (overwrite mode, rate 16, digest = first 8 words) ----
template <class F>
__global__ void __launch_bounds__(128) k_hash_rows_w24(const HashJob* __restrict__ jobs, uint32_t n_jobs,
                                                        const Poseidon2ConstsW* __restrict__ k) {
    uint32_t j = 0;
    while (j + 1 < n_jobs && blockIdx.x >= jobs[j + 1].cta_begin) j++;
    const HashJob job = jobs[j];
    uint32_t r = (blockIdx.x - job.cta_begin) * blockDim.x + threadIdx.x;
    if (r >= job.n_rows) return;
    uint32_t st[24];
#pragma unroll
    for (int i = 0; i < 24; i++) st[i] = 0;
    for (uint32_t c0 = 0; c0 < job.ncols; c0 += 16) {
#pragma unroll
        for (int q = 0; q < 16; q++)
            if (c0 + q < job.ncols) st[q] = __ldg(job.colptr[c0 + q] + r);
        poseidon2_permute_w<F, 24>(st, k);
    }
    uint4* o = reinterpret_cast<uint4*>(job.out + (size_t)r * 8);
    o[0] = make_uint4(st[0], st[1], st[2], st[3]);
    o[1] = make_uint4(st[4], st[5], st[6], st[7]);
}
template <class F>
__global__ void __launch_bounds__(128) k_hash_rows_rowmajor_w24(const uint32_t* __restrict__ data, uint32_t w, uint32_t n_rows,
                                                                 uint32_t* __restrict__ out, const Poseidon2ConstsW* __restrict__ k) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    uint32_t st[24];
#pragma unroll
    for (int i = 0; i < 24; i++) st[i] = 0;
    const uint32_t* row = data + (size_t)r * w;
    for (uint32_t c0 = 0; c0 < w; c0 += 16) {
#pragma unroll
        for (int q = 0; q < 16; q++)
            if (c0 + q < w) st[q] = row[c0 + q];
        poseidon2_permute_w<F, 24>(st, k);
    }
    uint4* o = reinterpret_cast<uint4*>(out + (size_t)r * 8);
    o[0] = make_uint4(st[0], st[1], st[2], st[3]);
    o[1] = make_uint4(st[4], st[5], st[6], st[7]);
}
template <class F, int W>
__global__ void k_permute_states_w(uint32_t* states, uint32_t n, const Poseidon2ConstsW* __restrict__ k) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t st[W];
#pragma unroll
    for (int q = 0; q < W; q++) st[q] = states[(size_t)i * W + q];
    poseidon2_permute_w<F, W>(st, k);
#pragma unroll
    for (int q = 0; q < W; q++) states[(size_t)i * W + q] = st[q];
}
// next[i] = compress(prev[2i], prev[2i+1]); with inj != nullptr additionally next[i] = compress(next[i], inj[i]) where inj
// holds the digests of the rows injected at this level.
template <class F>
__global__ void __launch_bounds__(128) k_compress(const uint32_t* __restrict__ prev, uint32_t* __restrict__ next, uint32_t n_next,
                                                   const uint32_t* __restrict__ inj) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_next) return;
    uint32_t st[16];
    const uint4* p = reinterpret_cast<const uint4*>(prev + (size_t)i * 16);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint4 v = p[q];
        st[4 * q] = v.x;
        st[4 * q + 1] = v.y;
        st[4 * q + 2] = v.z;
        st[4 * q + 3] = v.w;
    }
    poseidon2_permute<F>(st);
    if (inj) {
        const uint4* h = reinterpret_cast<const uint4*>(inj + (size_t)i * 8);
        uint4 h0 = h[0], h1 = h[1];
        st[8] = h0.x, st[9] = h0.y, st[10] = h0.z, st[11] = h0.w;
        st[12] = h1.x, st[13] = h1.y, st[14] = h1.z, st[15] = h1.w;
        poseidon2_permute<F>(st);
    }
    uint4* o = reinterpret_cast<uint4*>(next + (size_t)i * 8);
    o[0] = make_uint4(st[0], st[1], st[2], st[3]);
    o[1] = make_uint4(st[4], st[5], st[6], st[7]);
}
// ---- cooperative (16 lanes per permutation) Merkle kernel for the small levels ---------------------------------------
// Several consecutive Merkle levels in ONE launch. CTA b owns the subtree rooted at node b of the stage's last level:
// with K = n_levels it produces 2^(K-1-j) nodes of level first_level + j (j < K), keeps them in shared memory for the
// next level and writes every level to the tree in HBM (openings need all of them). One 16-lane group per node, so a
// level costs one cooperative-permutation latency (plus the injected row's sponge) instead of one launch; the levels of a
// tree with <= 2^13 nodes take ceil(levels/7) launches. Optionally the stage starts from the leaves (first_level == 0):
// the CTA first hashes its 2^K leaf rows of one row-major matrix (leaf_rows, the ExtensionMmcs rows of the FRI commit phase,
// recursion/src/pcs/mmcs.rs:434-441). Rows injected at a level arrive as digests (k_hash_rows).
// Level l (2^(log_max_h - l) digests) lives at digest offset 2^(log_max_h+1) - 2^(log_max_h-l+1).
constexpr uint32_t STAGE_MAX_LEVELS = 7;   // 2^(7-1) = 64 first-level nodes = the 64 lane groups of a 1024-thread CTA
constexpr uint32_t STAGE_DEFAULT_LEVELS = 5;   // measured best of 4..7 on the layer workload (more CTAs, no issue contention at the widest level)
struct MerkleStage {
    uint32_t* digests;
    uint32_t log_max_h;
    uint32_t first_level;   // first level produced by compression (>= 1); with leaves: levels 0 .. n_levels
    uint32_t n_levels;      // K compression levels (may be 0 with leaves)
    uint32_t with_leaves;   // hash the leaf level first
    const uint32_t* leaf_rows;            // row-major leaf matrix
    uint32_t leaf_w;                      // columns per leaf row
    const uint32_t* inj[STAGE_MAX_LEVELS];  // by (level - first_level): digests of the rows injected there, or nullptr
};
template <class F>
__global__ void __launch_bounds__(1024) k_merkle_stage(MerkleStage a, const Poseidon2Consts* __restrict__ gk) {
    __shared__ uint32_t buf[2][128 * 8];
    const uint32_t lane = threadIdx.x & 31u, l16 = lane & 15u;
    const P2Lane c = p2_lane_consts<F>(gk, l16);
    const uint32_t group = threadIdx.x >> 4, n_groups = blockDim.x >> 4;
    const uint32_t K = a.n_levels;
    auto level_ptr = [&](uint32_t l) {
        return a.digests + (((size_t)2 << a.log_max_h) - ((size_t)2 << (a.log_max_h - l))) * 8;
    };
    if (a.with_leaves) {
        const uint32_t per_cta = 1u << K;
        uint32_t* out = level_ptr(0);
        for (uint32_t base = 0; base < per_cta; base += n_groups) {
            const uint32_t loc = base + group;
            const bool live = loc < per_cta;
            if (__ballot_sync(0xffffffffu, live)) {
                const uint32_t r = blockIdx.x * per_cta + (live ? loc : 0);
                uint32_t x = 0;
                const uint32_t* row = a.leaf_rows + (size_t)r * a.leaf_w;
                for (uint32_t c0 = 0; c0 < a.leaf_w; c0 += 8) {
                    if (l16 < 8 && c0 + l16 < a.leaf_w) x = row[c0 + l16];
                    x = p2_coop_permute<F>(x, lane, c);
                }
                if (live && l16 < 8) {
                    out[(size_t)r * 8 + l16] = x;
                    buf[1][loc * 8 + l16] = x;
                }
            }
        }
        __syncthreads();
    }
    for (uint32_t j = 0; j < K; j++) {
        const uint32_t l = a.first_level + j;
        const uint32_t per_cta = 1u << (K - 1 - j);
        const uint32_t* prev = level_ptr(l - 1);
        uint32_t* next = level_ptr(l);
        const uint32_t* sprev = buf[(j + 1) & 1];
        uint32_t* snext = buf[j & 1];
        const bool from_smem = j > 0 || a.with_leaves;
        const uint32_t* inj = a.inj[j];
        for (uint32_t base = 0; base < per_cta; base += n_groups) {
            const uint32_t loc = base + group;
            const bool live = loc < per_cta;
            // a warp holds two groups: it runs when either is live (the dead half computes node 0 again and drops it)
            if (__ballot_sync(0xffffffffu, live)) {
                const uint32_t ll = live ? loc : 0;
                const uint32_t i = blockIdx.x * per_cta + ll;
                uint32_t x = from_smem ? sprev[ll * 16 + l16] : prev[(size_t)i * 16 + l16];  // left digest || right digest
                x = p2_coop_permute<F>(x, lane, c);
                if (inj) {
                    if (l16 >= 8) x = __ldg(inj + (size_t)i * 8 + (l16 - 8));  // lanes 8..15 take the injected digest
                    x = p2_coop_permute<F>(x, lane, c);
                }
                if (live && l16 < 8) {
                    next[(size_t)i * 8 + l16] = x;
                    snext[loc * 8 + l16] = x;
                }
            }
        }
        __syncthreads();
    }
}
template <class F>
__global__ void k_permute_states(uint32_t* states, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t st[16];
#pragma unroll
    for (int k = 0; k < 16; k++) st[k] = states[(size_t)i * 16 + k];
    poseidon2_permute<F>(st);
#pragma unroll
    for (int k = 0; k < 16; k++) states[(size_t)i * 16 + k] = st[k];
}

// K12: DuplexChallenger::grind. Thread w tests witness `base + w` (canonical): absorb pending||w, zero-fill the rate,
// tag the length into state[8], permute, sample = state[7] (pop from the back of the output buffer); accept when its low
// `bits` canonical bits are zero (recursion/src/challenger/circuit.rs:97-156,409-430). atomicMin keeps the smallest.
template <class F>
__global__ void __launch_bounds__(128) k_grind(const uint32_t* __restrict__ state, const uint32_t* __restrict__ pending,
                                                uint32_t n_pending, uint32_t bits, uint32_t base, uint32_t count,
                                                uint32_t* __restrict__ best) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t w = base + i;
    if (w >= F::P) return;
    uint32_t st[16];
#pragma unroll
    for (int k = 0; k < 16; k++) st[k] = state[k];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint32_t v = 0;
        if ((uint32_t)k < n_pending) v = pending[k];
        else if ((uint32_t)k == n_pending) v = to_monty<F>(w);
        st[k] = v;
    }
    st[8] = fadd<F>(st[8], to_monty<F>(n_pending + 1));
    poseidon2_permute<F>(st);
    uint32_t s = from_monty<F>(st[7]);
    if ((s & ((1u << bits) - 1)) == 0) atomicMin(best, w);
}

// ------------------------------------------------------------------------------------------------
// Constraint bytecode interpreter (format: include/p3r.h). One thread = one row; slots live in local memory.
// ------------------------------------------------------------------------------------------------
constexpr int MAX_B_SLOTS = 192;
constexpr int MAX_E_SLOTS = 24;

struct RowSrc {
    const uint32_t* main;   // column-major
    const uint32_t* prep;
    const uint32_t* perm;   // flattened base columns (4 per EF column)
    uint64_t main_cs, prep_cs, perm_cs;  // column strides
    uint32_t row[2];        // storage row of local / next
    const uint32_t* pub;
    uint32_t sel[3];
    const Ext4* chal;
    const Ext4* pval;
    const Ext4* econst;
};

// Sink: called for ASSERT_B/ASSERT_E/OUT_B.
template <class F, class Sink>
__device__ __forceinline__ void run_program(const uint4* __restrict__ insns, uint32_t n_insns, const RowSrc& rs, uint32_t wnr,
                                            Sink& sink) {
    uint32_t B[MAX_B_SLOTS];
    Ext4 E[MAX_E_SLOTS];
    for (uint32_t pc = 0; pc < n_insns; pc++) {
        uint4 in = __ldg(insns + pc);
        uint32_t op = in.x, d = in.y, a = in.z, b = in.w;
        switch (op) {
            case P3R_OP_B_MAIN: B[d] = __ldg(rs.main + (size_t)a * rs.main_cs + rs.row[b]); break;
            case P3R_OP_B_PREP: B[d] = __ldg(rs.prep + (size_t)a * rs.prep_cs + rs.row[b]); break;
            case P3R_OP_B_PUB: B[d] = __ldg(rs.pub + a); break;
            case P3R_OP_B_SEL: B[d] = rs.sel[a]; break;
            case P3R_OP_B_CONST: B[d] = a; break;
            case P3R_OP_B_ADD: B[d] = fadd<F>(B[a], B[b]); break;
            case P3R_OP_B_SUB: B[d] = fsub<F>(B[a], B[b]); break;
            case P3R_OP_B_MUL: B[d] = fmul<F>(B[a], B[b]); break;
            case P3R_OP_B_NEG: B[d] = fneg<F>(B[a]); break;
            case P3R_OP_E_PERM: {
                const uint32_t* p = rs.perm + (size_t)(4 * a) * rs.perm_cs + rs.row[b];
                Ext4 v;
                v.c[0] = __ldg(p);
                v.c[1] = __ldg(p + rs.perm_cs);
                v.c[2] = __ldg(p + 2 * rs.perm_cs);
                v.c[3] = __ldg(p + 3 * rs.perm_cs);
                E[d] = v;
                break;
            }
            case P3R_OP_E_CHAL: E[d] = rs.chal[a]; break;
            case P3R_OP_E_PVAL: E[d] = rs.pval[a]; break;
            case P3R_OP_E_CONST: E[d] = rs.econst[a]; break;
            case P3R_OP_E_FROMB: E[d] = ext_lift<F>(B[a]); break;
            case P3R_OP_E_ADD: E[d] = eadd<F>(E[a], E[b]); break;
            case P3R_OP_E_SUB: E[d] = esub<F>(E[a], E[b]); break;
            case P3R_OP_E_MUL: E[d] = emul<F>(E[a], E[b], wnr); break;
            case P3R_OP_E_NEG: E[d] = eneg<F>(E[a]); break;
            case P3R_OP_E_MULB: E[d] = emul_base<F>(E[a], B[b]); break;
            case P3R_OP_E_ADDB: E[d] = eadd_base<F>(E[a], B[b]); break;
            case P3R_OP_E_SUBB: E[d] = esub_base<F>(E[a], B[b]); break;
            case P3R_OP_ASSERT_B: sink.assert_b(d, B[a]); break;
            case P3R_OP_ASSERT_E: sink.assert_e(d, E[a]); break;
            case P3R_OP_OUT_B: sink.out_b(d, B[a]); break;
            default: break;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K6: LogUp permutation trace. Thread = trace row r (natural order, trace domain).
// Writes the fraction columns (perm EF column c+1, flattened) and the row sum; k_logup_scan turns row sums into the
// running accumulator column 0 and the terminal.
// ------------------------------------------------------------------------------------------------
constexpr int MAX_LK_OUTS = 160;
constexpr int MAX_INTERACTIONS = 32;

struct LogupArgs {
    const uint4* insns;
    uint32_t n_insns;
    const uint32_t* main;
    const uint32_t* prep;
    const uint32_t* pub;
    uint32_t log_n;
    const p3r_lookup* lookups;
    uint32_t n_lookups;
    const p3r_interaction* inter;
    const Ext4* chal;       // per lookup [prefix, beta]
    const Ext4* beta_pows;  // beta^k, k < 8
    uint32_t* perm;         // column-major flattened, (n_lookups+1)*4 columns of height n
    Ext4* rowsum;           // n entries
    uint32_t wnr;
    uint32_t cta_begin;     // multi-table launch (k_logup_rows): first CTA of this table
    Ext4* chunk_sum;        // scan scratch: one partial per SCAN_CHUNK rows
    Ext4* terminal;         // out: total of the table
    uint32_t scan_cta_begin;  // first CTA of this table in the two scan launches
};
struct OutSink {
    uint32_t* outs;
    __device__ __forceinline__ void assert_b(uint32_t, uint32_t) {}
    __device__ __forceinline__ void assert_e(uint32_t, const Ext4&) {}
    __device__ __forceinline__ void out_b(uint32_t i, uint32_t v) { outs[i] = v; }
};
// All tables with lookups in one launch (flat grid, largest tables first).
template <class F>
__global__ void __launch_bounds__(128) k_logup_rows(const LogupArgs* __restrict__ tables, uint32_t n_tables) {
    uint32_t jb = 0;
    while (jb + 1 < n_tables && blockIdx.x >= tables[jb + 1].cta_begin) jb++;
    const LogupArgs a = tables[jb];
    uint32_t n = 1u << a.log_n;
    uint32_t r = (blockIdx.x - a.cta_begin) * blockDim.x + threadIdx.x;
    if (r >= n) return;
    uint32_t outs[MAX_LK_OUTS];
    RowSrc rs;
    rs.main = a.main;
    rs.prep = a.prep;
    rs.perm = nullptr;
    rs.main_cs = rs.prep_cs = n;
    rs.perm_cs = 0;
    rs.row[0] = r;
    rs.row[1] = (r + 1) & (n - 1);
    rs.pub = a.pub;
    rs.sel[0] = (r == 0) ? F::R : 0;
    rs.sel[1] = (r == n - 1) ? F::R : 0;
    rs.sel[2] = (r != n - 1) ? F::R : 0;
    rs.chal = a.chal;
    rs.pval = nullptr;
    rs.econst = nullptr;
    OutSink sink{outs};
    run_program<F>(a.insns, a.n_insns, rs, a.wnr, sink);

    // denominators of all interactions, then one batched inversion (Montgomery's trick)
    Ext4 den[MAX_INTERACTIONS], pre[MAX_INTERACTIONS];
    uint32_t nint = 0;
    for (uint32_t c = 0; c < a.n_lookups; c++) {
        p3r_lookup l = a.lookups[c];
        Ext4 prefix = a.chal[2 * c];
        for (uint32_t j = 0; j < l.n_interactions; j++) {
            p3r_interaction it = a.inter[l.first_interaction + j];
            Ext4 d = prefix;
            for (uint32_t k = 0; k < it.n_elems; k++) d = eadd<F>(d, emul_base<F>(a.beta_pows[k], outs[it.elem_out_first + k]));
            den[nint++] = d;
        }
    }
    Ext4 acc = ext_one<F>();
    for (uint32_t i = 0; i < nint; i++) {
        pre[i] = acc;
        acc = emul<F>(acc, den[i], a.wnr);
    }
    Ext4 inv = einv<F>(acc, a.wnr);
    for (uint32_t i = nint; i-- > 0;) {
        Ext4 di = emul<F>(inv, pre[i], a.wnr);
        inv = emul<F>(inv, den[i], a.wnr);
        den[i] = di;  // now 1/den[i]
    }
    Ext4 total = ext_zero();
    uint32_t q = 0;
    for (uint32_t c = 0; c < a.n_lookups; c++) {
        p3r_lookup l = a.lookups[c];
        Ext4 frac = ext_zero();
        for (uint32_t j = 0; j < l.n_interactions; j++, q++) {
            p3r_interaction it = a.inter[l.first_interaction + j];
            frac = eadd<F>(frac, emul_base<F>(den[q], outs[it.mult_out]));
        }
#pragma unroll
        for (int k = 0; k < 4; k++) a.perm[(size_t)(4 * (c + 1) + k) * n + r] = frac.c[k];
        total = eadd<F>(total, frac);
    }
    a.rowsum[r] = total;
}
// Exclusive prefix sum of rowsum into perm columns 0..3 and the terminal, in two small launches:
// (1) one partial sum per 256-row chunk, (2) every chunk adds the partials before it and scans its own rows.
constexpr uint32_t SCAN_CHUNK = 256;
template <class F>
__device__ __forceinline__ Ext4 block_sum_256(Ext4 v, Ext4* red) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t x = v.c[k];
        for (int off = 16; off > 0; off >>= 1) x = fadd<F>(x, __shfl_down_sync(0xffffffffu, x, off));
        v.c[k] = x;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    Ext4 t = red[0];
    for (uint32_t q = 1; q < blockDim.x / 32; q++) t = eadd<F>(t, red[q]);
    __syncthreads();
    return t;
}
template <class F>
__global__ void __launch_bounds__(256) k_logup_chunk_sums(const LogupArgs* __restrict__ tables, uint32_t n_tables) {
    __shared__ Ext4 red[8];
    uint32_t jb = 0;
    while (jb + 1 < n_tables && blockIdx.x >= tables[jb + 1].scan_cta_begin) jb++;
    const Ext4* rowsum = tables[jb].rowsum;
    const uint32_t n = 1u << tables[jb].log_n, chunk = blockIdx.x - tables[jb].scan_cta_begin;
    uint32_t r = chunk * SCAN_CHUNK + threadIdx.x;
    Ext4 v = r < n ? rowsum[r] : ext_zero();
    Ext4 t = block_sum_256<F>(v, red);
    if (threadIdx.x == 0) tables[jb].chunk_sum[chunk] = t;
}
template <class F>
__global__ void __launch_bounds__(256) k_logup_scan_apply(const LogupArgs* __restrict__ tables, uint32_t n_tables) {
    __shared__ Ext4 red[8];
    __shared__ Ext4 buf[SCAN_CHUNK];
    uint32_t jb = 0;
    while (jb + 1 < n_tables && blockIdx.x >= tables[jb + 1].scan_cta_begin) jb++;
    const Ext4* rowsum = tables[jb].rowsum;
    const Ext4* chunk_sum = tables[jb].chunk_sum;
    uint32_t* perm = tables[jb].perm;
    Ext4* terminal = tables[jb].terminal;
    const uint32_t n = 1u << tables[jb].log_n, chunk = blockIdx.x - tables[jb].scan_cta_begin;
    // prefix over earlier chunks
    Ext4 pre = ext_zero();
    for (uint32_t c = threadIdx.x; c < chunk; c += blockDim.x) pre = eadd<F>(pre, chunk_sum[c]);
    pre = block_sum_256<F>(pre, red);
    uint32_t r = chunk * SCAN_CHUNK + threadIdx.x;
    Ext4 own = r < n ? rowsum[r] : ext_zero();
    buf[threadIdx.x] = own;
    __syncthreads();
    for (uint32_t off = 1; off < SCAN_CHUNK; off <<= 1) {  // Hillis-Steele inclusive scan
        Ext4 v = buf[threadIdx.x];
        if (threadIdx.x >= off) v = eadd<F>(v, buf[threadIdx.x - off]);
        __syncthreads();
        buf[threadIdx.x] = v;
        __syncthreads();
    }
    Ext4 incl = eadd<F>(pre, buf[threadIdx.x]);
    Ext4 excl = esub<F>(incl, own);
    if (r < n) {
#pragma unroll
        for (int k = 0; k < 4; k++) perm[(size_t)k * n + r] = excl.c[k];
        if (r == n - 1) *terminal = incl;
    }
}

// ------------------------------------------------------------------------------------------------
// K7: quotient evaluation. Thread = storage row s of the quotient domain (first n*qc rows of the bit-reversed LDE);
// natural index i = bitrev(s), next row = i + qc. Selectors follow p3's selectors_on_coset (trace domain shift 1):
// Z_H(x) = x^n - 1, is_first = Z_H/(x-1), is_last = Z_H/(x - g^-1), is_transition = x - g^-1, inv_vanishing = 1/Z_H.
// Folding: sum_k alpha^{N-1-k} * c_k  ==  Horner acc = acc*alpha + c (recursion/src/traits/air.rs:170-181).
// Output: natural-order chunk matrices: chunk (i mod qc), row (i div qc), 4 base columns each (split_evals).
// ------------------------------------------------------------------------------------------------
struct QuotientArgs {
    const uint4* insns;
    uint32_t n_insns;
    const uint32_t* main;
    const uint32_t* prep;
    const uint32_t* perm;
    const uint32_t* pub;
    uint32_t log_n, log_qc, log_blowup;
    const uint32_t* sel;        // 3 arrays of NQ (storage order): is_first, is_last, is_transition
    const uint32_t* inv_van;    // qc entries: 1/Z_H on coset class (i mod qc)
    const Ext4* chal;
    const Ext4* pval;
    const Ext4* econst;
    const Ext4* alpha_pows;     // alpha^{N-1-k} at index k
    uint32_t* chunks;           // qc matrices, each 4 columns x n rows, column-major: [(c*4 + k)*n + r]
    uint32_t wnr;
};
template <class F>
struct FoldSink {
    const Ext4* ap;
    uint32_t wnr;
    Ext4 acc;
    __device__ __forceinline__ void assert_b(uint32_t k, uint32_t v) { acc = eadd<F>(acc, emul_base<F>(ap[k], v)); }
    __device__ __forceinline__ void assert_e(uint32_t k, const Ext4& v) { acc = eadd<F>(acc, emul<F>(ap[k], v, wnr)); }
    __device__ __forceinline__ void out_b(uint32_t, uint32_t) {}
};
template <class F>
__global__ void __launch_bounds__(128) k_quotient(QuotientArgs a) {
    const uint32_t lq = a.log_n + a.log_qc, NQ = 1u << lq, n = 1u << a.log_n;
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= NQ) return;
    uint32_t i = bitrev32(s, lq);
    uint32_t inext = (i + (1u << a.log_qc)) & (NQ - 1);
    RowSrc rs;
    rs.main = a.main;
    rs.prep = a.prep;
    rs.perm = a.perm;
    rs.main_cs = rs.prep_cs = rs.perm_cs = (uint64_t)n << a.log_blowup;
    rs.row[0] = s;
    rs.row[1] = bitrev32(inext, lq);
    rs.pub = a.pub;
    rs.sel[0] = a.sel[s];
    rs.sel[1] = a.sel[NQ + s];
    rs.sel[2] = a.sel[2 * NQ + s];
    rs.chal = a.chal;
    rs.pval = a.pval;
    rs.econst = a.econst;
    FoldSink<F> sink{a.alpha_pows, a.wnr, ext_zero()};
    run_program<F>(a.insns, a.n_insns, rs, a.wnr, sink);
    Ext4 q = emul_base<F>(sink.acc, a.inv_van[i & ((1u << a.log_qc) - 1)]);
    uint32_t c = i & ((1u << a.log_qc) - 1), r = i >> a.log_qc;
#pragma unroll
    for (int k = 0; k < 4; k++) a.chunks[((size_t)c * 4 + k) * n + r] = q.c[k];
}
// Interpreter with constraint groups: the program is cut into QG_GROUPS sub-programs at preparation time (p3r.cu slice_program:
// the constraints of a group plus, by backward liveness over the slot-allocated code, exactly the instructions they need), a
// CTA is 32 rows x QG_GROUPS warps, warp g folds group g and the partial sums meet in shared memory — the schedule of the
// generated kernels (k_quotient_spec_*) for programs that have none (other packings, conventions, the wide uni-stark table,
// whose 13 000-instruction program took 4.1 ms at one thread per row whatever the number of rows).
constexpr int QG_GROUPS = 8;
struct QuotientGroups {
    const uint4* insns;             // the sub-programs back to back
    uint32_t off[QG_GROUPS + 1];    // instruction range of group g: [off[g], off[g + 1])
};
template <class F>
__global__ void __launch_bounds__(32 * QG_GROUPS) k_quotient_grouped(QuotientArgs a, QuotientGroups gr) {
    __shared__ Ext4 part[QG_GROUPS][32];
    const uint32_t lq = a.log_n + a.log_qc, NQ = 1u << lq, n = 1u << a.log_n;
    const uint32_t g = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t s_raw = blockIdx.x * 32 + lane;
    const uint32_t s = s_raw < NQ ? s_raw : NQ - 1;   // out-of-range lanes recompute the last row and drop it
    const uint32_t i = bitrev32(s, lq);
    const uint32_t inext = (i + (1u << a.log_qc)) & (NQ - 1);
    RowSrc rs;
    rs.main = a.main;
    rs.prep = a.prep;
    rs.perm = a.perm;
    rs.main_cs = rs.prep_cs = rs.perm_cs = (uint64_t)n << a.log_blowup;
    rs.row[0] = s;
    rs.row[1] = bitrev32(inext, lq);
    rs.pub = a.pub;
    rs.sel[0] = a.sel[s];
    rs.sel[1] = a.sel[NQ + s];
    rs.sel[2] = a.sel[2 * NQ + s];
    rs.chal = a.chal;
    rs.pval = a.pval;
    rs.econst = a.econst;
    FoldSink<F> sink{a.alpha_pows, a.wnr, ext_zero()};
    run_program<F>(gr.insns + gr.off[g], gr.off[g + 1] - gr.off[g], rs, a.wnr, sink);
    part[g][lane] = sink.acc;
    __syncthreads();
    if (g == 0 && s_raw < NQ) {
        Ext4 tot = part[0][lane];
#pragma unroll
        for (int k = 1; k < QG_GROUPS; k++) tot = eadd<F>(tot, part[k][lane]);
        const Ext4 q = emul_base<F>(tot, a.inv_van[i & ((1u << a.log_qc) - 1)]);
        const uint32_t c = i & ((1u << a.log_qc) - 1), r = i >> a.log_qc;
#pragma unroll
        for (int k = 0; k < 4; k++) a.chunks[((size_t)c * 4 + k) * n + r] = q.c[k];
    }
}
// Selector arrays for (log_n, log_qc) in storage order + inv_vanishing per coset class.
template <class F>
__global__ void k_selectors(uint32_t* sel, uint32_t* inv_van, uint32_t log_n, uint32_t log_qc, uint32_t gen,
                            const uint32_t* tw, uint32_t logT) {
    const uint32_t lq = log_n + log_qc, NQ = 1u << lq;
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= NQ) return;
    uint32_t i = bitrev32(s, lq);
    uint32_t x = fmul<F>(gen, root_pow<F>(tw, logT, (uint64_t)i << (logT - lq)));
    uint32_t ginv = root_pow<F>(tw, logT, ((uint64_t)1 << logT) - ((uint64_t)1 << (logT - log_n)));
    uint32_t xn = x;
    for (uint32_t k = 0; k < log_n; k++) xn = fmul<F>(xn, xn);
    uint32_t z = fsub<F>(xn, F::R);
    sel[s] = fmul<F>(z, finv<F>(fsub<F>(x, F::R)));
    sel[NQ + s] = fmul<F>(z, finv<F>(fsub<F>(x, ginv)));
    sel[2 * NQ + s] = fsub<F>(x, ginv);
    if (i < (1u << log_qc)) inv_van[i] = finv<F>(z);
}
// out[k] = alpha^{n-1-k} for up to 8 tables in one launch (blockIdx.y = table): every thread raises alpha to its own power.
struct PowDescJobs {
    Ext4* out[8];
    uint32_t n[8];
};
template <class F>
__global__ void __launch_bounds__(128) k_ext_powers_desc(PowDescJobs jobs, Ext4 alpha, uint32_t wnr) {
    const uint32_t n = jobs.n[blockIdx.y];
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    jobs.out[blockIdx.y][k] = epow<F>(alpha, n - 1 - k, wnr);
}
template <class F>
__global__ void k_ext_powers_asc(Ext4* out, uint32_t n, Ext4 alpha, uint32_t wnr) {
    // out[k] = alpha^k, blocked: thread t computes alpha^(t*8) then 8 steps.
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t lo = t * 8;
    if (lo >= n) return;
    Ext4 acc = epow<F>(alpha, lo, wnr);
    for (uint32_t k = lo; k < min(lo + 8, n); k++) {
        out[k] = acc;
        acc = emul<F>(acc, alpha, wnr);
    }
}

// ------------------------------------------------------------------------------------------------
// K8: out-of-domain openings by barycentric evaluation over the source evaluations (n points of in_shift*H_n, natural
// order): p(z) = ((u^n - 1)/n) * sum_i p_i * w^i / (u - w^i), u = z / in_shift.
// ------------------------------------------------------------------------------------------------
struct WeightJob {
    Ext4 u;          // z / in_shift
    Ext4 scale;      // (u^n - 1) / n (host-computed)
    uint32_t log_n;
    uint32_t offset; // into the weights buffer (Ext4 units)
};
// weights[i] = scale * w^i / (u - w^i). Each thread produces 4 weights (rows i, i + n/4, ...) and shares ONE extension
// inversion between them (Montgomery's trick): the inversion's base-field Fermat power is most of the cost.
template <class F>
__global__ void __launch_bounds__(256) k_bary_weights(const WeightJob* __restrict__ jobs, Ext4* __restrict__ weights,
                                                       const uint32_t* tw, uint32_t logT, uint32_t wnr) {
    const WeightJob j = jobs[blockIdx.y];
    const uint32_t n = 1u << j.log_n;
    const uint32_t q = n >= 4 ? n / 4 : n, per = n >= 4 ? 4 : 1;
    const uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (i0 >= q) return;
    uint32_t wi[4];
    Ext4 den[4], pre[4];
    Ext4 acc = ext_one<F>();
    for (uint32_t t = 0; t < per; t++) {
        wi[t] = root_pow<F>(tw, logT, (uint64_t)(i0 + t * q) << (logT - j.log_n));
        den[t] = esub_base<F>(j.u, wi[t]);
        pre[t] = acc;
        acc = emul<F>(acc, den[t], wnr);
    }
    Ext4 inv = einv<F>(acc, wnr);
    for (uint32_t t = per; t-- > 0;) {
        Ext4 di = emul<F>(inv, pre[t], wnr);  // 1 / den[t]
        inv = emul<F>(inv, den[t], wnr);
        weights[j.offset + i0 + t * q] = emul<F>(emul_base<F>(j.scale, wi[t]), di, wnr);
    }
}
struct DotJob {
    const uint32_t* mat;   // column-major, height 2^log_n
    uint32_t log_n, width;
    uint32_t w_offset;     // weights
    uint32_t out_offset;   // into opened values buffer (Ext4 units), width entries
};
constexpr uint32_t DOT_ROWS = 8192;  // rows per CTA
constexpr uint32_t DOT_COLS = 4;     // columns per CTA (8 needs 158 registers: one CTA per SM)
struct DotTile {           // one CTA of k_bary_dot: rows [chunk*DOT_ROWS, ..) x columns [c0, c0 + DOT_COLS) of a job
    uint32_t job, chunk, c0;
};
// partial[job][chunk][col]. Every thread walks its rows once, keeps the weight in registers and accumulates the DOT_COLS
// columns in 64-bit sums of four Montgomery products (one reduction per four rows).
template <class F>
__global__ void __launch_bounds__(256, 3) k_bary_dot(const DotJob* __restrict__ jobs, const DotTile* __restrict__ tiles,
                                                   const Ext4* __restrict__ weights, Ext4* __restrict__ partial,
                                                   uint32_t max_chunks, uint32_t max_width) {
    const DotTile tl = tiles[blockIdx.x];
    const DotJob j = jobs[tl.job];
    const uint32_t n = 1u << j.log_n;
    const uint32_t r0 = tl.chunk * DOT_ROWS, r1 = min(r0 + DOT_ROWS, n);
    const uint32_t nc = min(DOT_COLS, j.width - tl.c0);
    __shared__ Ext4 red[8][DOT_COLS];
    const Ext4* w = weights + j.w_offset;
    const uint32_t* col0 = j.mat + (size_t)tl.c0 * n;
    Ext4 acc[DOT_COLS];
    uint64_t a[DOT_COLS][4];
#pragma unroll
    for (int c = 0; c < (int)DOT_COLS; c++) {
        acc[c] = ext_zero();
        a[c][0] = a[c][1] = a[c][2] = a[c][3] = 0;
    }
    uint32_t cnt = 0;
    // columns past the job's width are clamped to its last column (computed, never stored): no predicated accumulates
    const uint32_t* colp[DOT_COLS];
#pragma unroll
    for (int c = 0; c < (int)DOT_COLS; c++) colp[c] = col0 + (size_t)min((uint32_t)c, nc - 1) * n;
    const uint4* w4 = reinterpret_cast<const uint4*>(w);   // weights are 16-byte aligned (arena)
    for (uint32_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
        const uint4 wr = __ldg(w4 + r);
#pragma unroll
        for (int c = 0; c < (int)DOT_COLS; c++) {
            const uint32_t v = __ldg(colp[c] + r);
            a[c][0] += (uint64_t)v * wr.x;
            a[c][1] += (uint64_t)v * wr.y;
            a[c][2] += (uint64_t)v * wr.z;
            a[c][3] += (uint64_t)v * wr.w;
        }
        if (++cnt == 4) {  // 4 products < 2^64
#pragma unroll
            for (int c = 0; c < (int)DOT_COLS; c++) {
                acc[c] = eadd<F>(acc[c], Ext4{{fred64<F>(a[c][0]), fred64<F>(a[c][1]), fred64<F>(a[c][2]), fred64<F>(a[c][3])}});
                a[c][0] = a[c][1] = a[c][2] = a[c][3] = 0;
            }
            cnt = 0;
        }
    }
#pragma unroll
    for (int c = 0; c < (int)DOT_COLS; c++) {
        acc[c] = eadd<F>(acc[c], Ext4{{fred64<F>(a[c][0]), fred64<F>(a[c][1]), fred64<F>(a[c][2]), fred64<F>(a[c][3])}});
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t v = acc[c].c[k];
            for (int off = 16; off > 0; off >>= 1) v = fadd<F>(v, __shfl_down_sync(0xffffffffu, v, off));
            acc[c].c[k] = v;
        }
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][c] = acc[c];
    }
    __syncthreads();
    if (threadIdx.x < nc) {
        Ext4 t = red[0][threadIdx.x];
        for (uint32_t q = 1; q < blockDim.x / 32; q++) t = eadd<F>(t, red[q][threadIdx.x]);
        partial[((size_t)tl.job * max_chunks + tl.chunk) * max_width + tl.c0 + threadIdx.x] = t;
    }
}
template <class F>
__global__ void k_bary_reduce(const DotJob* __restrict__ jobs, const Ext4* __restrict__ partial, Ext4* __restrict__ opened,
                              uint32_t max_chunks, uint32_t max_width) {
    DotJob j = jobs[blockIdx.y];
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= j.width) return;
    uint32_t n = 1u << j.log_n;
    uint32_t chunks = (n + DOT_ROWS - 1) / DOT_ROWS;
    Ext4 t = ext_zero();
    for (uint32_t q = 0; q < chunks; q++) t = eadd<F>(t, partial[((size_t)blockIdx.y * max_chunks + q) * max_width + c]);
    opened[j.out_offset + c] = t;
}

// ------------------------------------------------------------------------------------------------
// K9: reduced openings per LDE height. For storage row s at log-height h (x = GEN * w_h^{bitrev(s)}):
//   ro[s] += sum over (matrix m, point j):  alpha^{off_{m,j}} * (P_{m,j} - R_m(s)) / (z_j - x),
//   R_m(s) = sum_k alpha^k * lde_m[k][s],   P_{m,j} = sum_k alpha^k * opened_{m,j}[k].
// ------------------------------------------------------------------------------------------------
struct RoMat {
    const uint32_t* lde;     // column-major, height 2^log_h
    uint32_t width;
    uint32_t n_points;       // 1 or 2 (point 0 = zeta, point 1 = zeta*g)
    uint32_t opened_off[2];  // Ext4 offset of the opened values of each point
    uint32_t alpha_off[2];   // exponent offset alpha^{off}
};
template <class F>
__global__ void __launch_bounds__(128) k_ro_prepare(const RoMat* __restrict__ mats, uint32_t n_mats, const Ext4* __restrict__ opened,
                                                     const Ext4* __restrict__ apow, Ext4* __restrict__ coef /* [mat][2]: alpha^off * P */,
                                                     uint32_t wnr) {
    // one warp per (matrix, point): lanes stride over the columns, then a shuffle reduction
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (t >= n_mats * 2) return;
    const RoMat m = mats[t >> 1];
    const uint32_t j = t & 1;
    if (j >= m.n_points) return;
    Ext4 acc = ext_zero();
    for (uint32_t k = lane; k < m.width; k += 32) acc = eadd<F>(acc, emul<F>(apow[k], opened[m.opened_off[j] + k], wnr));
#pragma unroll
    for (int c = 0; c < 4; c++) {
        uint32_t v = acc.c[c];
        for (int off = 16; off > 0; off >>= 1) v = fadd<F>(v, __shfl_down_sync(0xffffffffu, v, off));
        acc.c[c] = v;
    }
    if (lane == 0) coef[t] = emul<F>(acc, apow[m.alpha_off[j]], wnr);
}
struct RoArgs {
    const RoMat* mats;
    uint32_t n_mats;
    uint32_t log_h;
    const Ext4* apow;       // alpha^k
    const Ext4* coef;       // alpha^off * P per (mat, point)
    Ext4 z[2];              // zeta, zeta*g for this height
    uint32_t gen;
    const uint32_t* tw;
    uint32_t logT;
    Ext4* ro;               // 2^log_h entries (overwritten)
    uint32_t wnr;
    uint32_t max_width;     // widest matrix of this height (alpha powers staged in shared memory)
    uint32_t cta_begin;     // multi-height launch: first CTA of this height
};
// All heights in one launch (flat grid, jobs sorted by decreasing height). Per row: R_m = sum_k alpha^k * lde_m[k][s]
// accumulated as 64-bit sums of four Montgomery products (alpha powers from shared memory, four independent loads in flight),
// and one extension inversion shared by 1/(zeta - x) and 1/(zeta*g - x).
template <class F>
__global__ void __launch_bounds__(128) k_reduced_openings(const RoArgs* __restrict__ jobs, uint32_t n_jobs) {
    extern __shared__ __align__(16) uint4 sap[];   // alpha^k, one 16-byte load each
    uint32_t jb = 0;
    while (jb + 1 < n_jobs && blockIdx.x >= jobs[jb + 1].cta_begin) jb++;
    const RoArgs& a = jobs[jb];
    for (uint32_t k = threadIdx.x; k < a.max_width; k += blockDim.x) {
        const Ext4 ap = a.apow[k];
        sap[k] = make_uint4(ap.c[0], ap.c[1], ap.c[2], ap.c[3]);
    }
    __syncthreads();
    const uint32_t N = 1u << a.log_h;
    const uint32_t s = (blockIdx.x - a.cta_begin) * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const uint32_t wnr = a.wnr;
    uint32_t x = fmul<F>(a.gen, root_pow<F>(a.tw, a.logT, (uint64_t)bitrev32(s, a.log_h) << (a.logT - a.log_h)));
    const Ext4 d0 = esub_base<F>(a.z[0], x), d1 = esub_base<F>(a.z[1], x);
    const Ext4 inv01 = einv<F>(emul<F>(d0, d1, wnr), wnr);
    const Ext4 inv0 = emul<F>(inv01, d1, wnr), inv1 = emul<F>(inv01, d0, wnr);
    Ext4 acc = ext_zero();
    for (uint32_t mi = 0; mi < a.n_mats; mi++) {
        const RoMat m = a.mats[mi];
        Ext4 R = ext_zero();
        const uint32_t* p = m.lde + s;
        const uint32_t* q = p;   // running column pointer (one 64-bit add per column instead of a shift + add + scale)
        uint32_t k = 0;
        for (; k + 4 <= m.width; k += 4, q += 4 * (size_t)N) {
            const uint32_t v0 = __ldg(q), v1 = __ldg(q + N), v2 = __ldg(q + 2 * (size_t)N), v3 = __ldg(q + 3 * (size_t)N);
            const uint4 a0 = sap[k], a1 = sap[k + 1], a2 = sap[k + 2], a3 = sap[k + 3];
            Ext4 t;
            t.c[0] = fred64<F>((uint64_t)v0 * a0.x + (uint64_t)v1 * a1.x + (uint64_t)v2 * a2.x + (uint64_t)v3 * a3.x);
            t.c[1] = fred64<F>((uint64_t)v0 * a0.y + (uint64_t)v1 * a1.y + (uint64_t)v2 * a2.y + (uint64_t)v3 * a3.y);
            t.c[2] = fred64<F>((uint64_t)v0 * a0.z + (uint64_t)v1 * a1.z + (uint64_t)v2 * a2.z + (uint64_t)v3 * a3.z);
            t.c[3] = fred64<F>((uint64_t)v0 * a0.w + (uint64_t)v1 * a1.w + (uint64_t)v2 * a2.w + (uint64_t)v3 * a3.w);
            R = eadd<F>(R, t);
        }
        for (; k < m.width; k++) {
            const uint4 ak = sap[k];
            R = eadd<F>(R, emul_base<F>(Ext4{{ak.x, ak.y, ak.z, ak.w}}, __ldg(p + (size_t)k * N)));
        }
        for (uint32_t j = 0; j < m.n_points; j++) {
            // alpha^off * (P - R) = coef - alpha^off * R
            Ext4 t = esub<F>(a.coef[2 * mi + j], emul<F>(a.apow[m.alpha_off[j]], R, wnr));
            acc = eadd<F>(acc, emul<F>(t, j ? inv1 : inv0, wnr));
        }
    }
    a.ro[s] = acc;
}

// ------------------------------------------------------------------------------------------------
// K10: FRI fold of a bit-reversed EF vector by arity 2^k (k sequential arity-2 folds with beta, beta^2, ...) and roll-in
// of the reduced opening of the folded height (recursion/src/pcs/fri/verifier.rs:564-585,772-777).
// fold(e0,e1) at x0 = w_L^{bitrev(i)}: (e0+e1)/2 + beta*(e0-e1)/(2*x0).
// ------------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128) k_fri_fold(const Ext4* __restrict__ in, Ext4* __restrict__ out, uint32_t log_len,
                                                   uint32_t log_arity, Ext4 beta, const Ext4* __restrict__ beta_dev,
                                                   const Ext4* __restrict__ roll, uint32_t inv2, const uint32_t* tw, uint32_t logT,
                                                   uint32_t wnr) {
    uint32_t out_len = 1u << (log_len - log_arity);
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= out_len) return;
    if (beta_dev) beta = *beta_dev;   // challenge sampled on the device (k_fri_round_transcript)
    Ext4 v[16];
    uint32_t arity = 1u << log_arity;
    for (uint32_t j = 0; j < arity; j++) v[j] = in[(size_t)i * arity + j];
    Ext4 b = beta;
    for (uint32_t step = 0; step < log_arity; step++) {
        uint32_t lg = log_len - step;          // log length of the vector being folded
        uint32_t half = arity >> (step + 1);   // outputs this thread produces at this step
        for (uint32_t j = 0; j < half; j++) {
            uint32_t gi = i * half + j;        // index in the folded vector (length 2^(lg-1))
            // 1/x0 = w_{2^lg}^{-e}, e = bitrev(gi, lg-1). Looked up directly, the bit-reversed index makes every lane of a warp
            // read a different cache line of the twiddle table (the kernel ran at 8-35 % of HBM on that gather alone). Split
            // instead: e = (rev5(gi & 31) << (lg-6)) + bitrev(gi >> 5, lg-6), so w^-e = w_64^-rev5(gi & 31) * w_{2^lg}^-bitrev(gi >> 5):
            // the first factor comes from 32 fixed table entries, the second is one address per 32 consecutive outputs.
            uint32_t xinv;
            if (lg >= 7) {
                const uint32_t f1 = root_pow<F>(tw, logT, (uint64_t)(64u - (__brev(gi & 31u) >> 27)) << (logT - 6));
                const uint32_t e_hi = bitrev32(gi >> 5, lg - 6);
                const uint32_t f2 = root_pow<F>(tw, logT, (((uint64_t)1 << lg) - e_hi) << (logT - lg));
                xinv = fmul<F>(f1, f2);
            } else {
                uint32_t e = bitrev32(gi, lg - 1);
                xinv = root_pow<F>(tw, logT, (((uint64_t)1 << lg) - e) << (logT - lg));
            }
            Ext4 e0 = v[2 * j], e1 = v[2 * j + 1];
            Ext4 sum = eadd<F>(e0, e1);
            Ext4 dif = emul_base<F>(esub<F>(e0, e1), xinv);
            v[j] = emul_base<F>(eadd<F>(sum, emul<F>(b, dif, wnr)), inv2);
        }
        b = emul<F>(b, b, wnr);
    }
    Ext4 r = v[0];
    if (roll) r = eadd<F>(r, emul<F>(b, roll[i], wnr));
    out[i] = r;
}
// Final polynomial: coefficients of the degree < 2^log_fpl interpolant of the folded evaluations (first 2^log_fpl storage
// entries = the size-2^log_fpl subgroup in bit-reversed order). Naive O(n^2) inverse DFT: n <= 64.
template <class F>
__global__ void k_final_poly(const Ext4* __restrict__ folded, Ext4* __restrict__ coeffs, uint32_t log_fpl, const uint32_t* tw,
                             uint32_t logT) {
    uint32_t n = 1u << log_fpl;
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Ext4 acc = ext_zero();
    for (uint32_t i = 0; i < n; i++) {
        Ext4 v = folded[bitrev32(i, log_fpl)];  // evaluation at w_n^i
        uint64_t e = ((uint64_t)n - ((uint64_t)i * k) % n) % n;
        acc = eadd<F>(acc, emul_base<F>(v, root_pow<F>(tw, logT, e << (logT - log_fpl))));
    }
    coeffs[k] = emul_base<F>(acc, finv<F>(to_monty<F>(n)));
}

// DuplexChallenger on the device for the FRI commit rounds of the one-shot prover (same semantics as the
// host HostChallenger: overwrite mode, zero-fill, length tag in state[8], samples popped from the back). One warp: lanes
// 0..15 hold the sponge state for the cooperative permutation. Per round it observes the round's cap and samples beta, so the
// host does not have to synchronise between rounds; the host replays the same operations afterwards on its own challenger.
struct DevChallenger {
    uint32_t st[16];
    uint32_t in[8], out[8];
    uint32_t n_in, n_out;
};
template <class F>
__global__ void __launch_bounds__(32) k_fri_round_transcript(DevChallenger* __restrict__ ch, const uint32_t* __restrict__ cap,
                                                             uint32_t n_cap_words, Ext4* __restrict__ beta_out,
                                                             const Poseidon2Consts* __restrict__ gk) {
    __shared__ DevChallenger s;
    const uint32_t lane = threadIdx.x, l16 = lane & 15u;
    const P2Lane c = p2_lane_consts<F>(gk, l16);
    for (uint32_t i = lane; i < sizeof(DevChallenger) / 4; i += 32) reinterpret_cast<uint32_t*>(&s)[i] = reinterpret_cast<const uint32_t*>(ch)[i];
    __syncwarp();
    auto duplex = [&]() {
        // every lane runs the permutation; lanes 0..15 carry the state
        uint32_t x = s.st[l16];
        const uint32_t n_in = s.n_in;
        if (l16 < n_in) x = s.in[l16];
        else if (n_in > 0 && l16 < 8) x = 0;
        if (n_in > 0 && l16 == 8) x = fadd<F>(x, to_monty<F>(n_in));
        x = p2_coop_permute<F>(x, lane, c);
        __syncwarp();
        if (lane < 16) s.st[lane] = x;
        if (lane < 8) s.out[lane] = x;
        if (lane == 0) {
            s.n_in = 0;
            s.n_out = 8;
        }
        __syncwarp();
    };
    for (uint32_t w = 0; w < n_cap_words; w++) {   // observe
        if (lane == 0) {
            s.n_out = 0;
            s.in[s.n_in] = cap[w];
            s.n_in = s.n_in + 1;
        }
        __syncwarp();
        if (s.n_in == 8) duplex();
    }
    Ext4 beta;
    for (int k = 0; k < 4; k++) {                  // sample_ext
        if (s.n_in > 0 || s.n_out == 0) duplex();
        beta.c[k] = s.out[s.n_out - 1];
        __syncwarp();
        if (lane == 0) s.n_out = s.n_out - 1;
        __syncwarp();
    }
    if (lane == 0) *beta_out = beta;
    for (uint32_t i = lane; i < sizeof(DevChallenger) / 4; i += 32) reinterpret_cast<uint32_t*>(ch)[i] = reinterpret_cast<const uint32_t*>(&s)[i];
}

// ------------------------------------------------------------------------------------------------
// K11: query openings. Static gather plan per session: for every (query-independent) segment either
//   kind 0: matrix row  — `width` words from a column-major matrix at row (index >> shift)
//   kind 1: merkle path — `depth` sibling digests of tree `layers` for leaf (index >> shift)
//   kind 2: FRI siblings — (arity-1) EF values of row (index >> shift >> log_arity) of a row-major EF matrix
// ------------------------------------------------------------------------------------------------
struct GatherSeg {
    uint32_t kind;
    const uint32_t* base;     // matrix / digest layers / EF vector
    uint32_t log_h;           // matrix: log height; path: log leaves; fri: log length of the vector
    uint32_t width;           // matrix width; path depth; fri: log_arity
    uint32_t shift;           // index >> shift
    uint32_t out_off;         // word offset within one query's blob
};
static __global__ void __launch_bounds__(256) k_query_gather(const GatherSeg* __restrict__ segs, uint32_t n_segs,
                                                       const uint32_t* __restrict__ indices, uint32_t words_per_query,
                                                       uint32_t* __restrict__ out) {
    uint32_t q = blockIdx.x;
    uint32_t index = indices[q];
    uint32_t* o = out + (size_t)q * words_per_query;
    for (uint32_t si = blockIdx.y; si < n_segs; si += gridDim.y) {
        GatherSeg g = segs[si];
        uint32_t idx = index >> g.shift;
        if (g.kind == 0) {
            size_t H = (size_t)1 << g.log_h;
            for (uint32_t c = threadIdx.x; c < g.width; c += blockDim.x) o[g.out_off + c] = g.base[(size_t)c * H + idx];
        } else if (g.kind == 1) {
            // layers are stored back to back: level l (2^(log_h-l) digests) at digest offset 2^(log_h+1) - 2^(log_h-l+1)
            for (uint32_t t = threadIdx.x; t < g.width * 8; t += blockDim.x) {
                uint32_t l = t >> 3, k = t & 7;
                size_t lvl_off = ((size_t)2 << g.log_h) - ((size_t)2 << (g.log_h - l));
                size_t node = (idx >> l) ^ 1;
                o[g.out_off + t] = g.base[(lvl_off + node) * 8 + k];
            }
        } else {
            uint32_t arity = 1u << g.width;
            uint32_t row = idx >> g.width, own = idx & (arity - 1);
            for (uint32_t t = threadIdx.x; t < (arity - 1) * 4; t += blockDim.x) {
                uint32_t j = t >> 2, k = t & 3;
                uint32_t src = j < own ? j : j + 1;
                o[g.out_off + t] = g.base[((size_t)row * arity + src) * 4 + k];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3: Poseidon2 table fill (Poseidon2CircuitAir::generate_trace_rows, /root/reference poseidon2-circuit-air/src/air.rs:280-520
// -> p3_poseidon2_air::generate_trace_rows_for_perm [P3-EXT]). One thread per table row: runs the permutation and stores
// every committed round value straight into the column-major device trace:
//   inputs[16] | 4 x (sbox regs[16*R], post[16]) | P x (sbox reg[R], post_sbox) | 4 x (sbox regs, post) | mmcs_bit | mmcs_index_sum
// with R = 1 register (x^3) for the degree-7 S-box, 0 for degree 3. Rows >= n_ops are the padding rows: real permutations of
// the zero state (batch_stark_prover/poseidon2.rs:1121-1140). The index accumulator (pass 1 of the reference, sequential) is
// recomputed per row by walking back to the start of its Merkle chain: acc = 2*acc + bit on chained Merkle rows.
// ------------------------------------------------------------------------------------------------
struct P2FillArgs {
    const uint32_t* inputs;        // n_ops x 16, row-major, Montgomery
    const uint8_t* mmcs_bit;       // n_ops
    const uint32_t* idx_sum;       // n_ops: op.mmcs_index_sum (used where the accumulator restarts), Montgomery
    const uint32_t* new_start;     // preprocessed column (device, natural order, height H): non-zero = chain start
    const uint32_t* merkle_path;   // preprocessed column
    uint32_t n_ops, log_h;
    uint32_t* out;                 // column-major main trace, height H
};
template <class F>
__global__ void __launch_bounds__(128) k_poseidon2_table_fill(P2FillArgs a) {
    const uint32_t H = 1u << a.log_h;
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= H) return;
    const Poseidon2Consts& k = c_p2[FieldId<F>::value];
    constexpr int R = (F::SBOX == 7) ? 1 : 0;
    uint32_t s[16];
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = r < a.n_ops ? a.inputs[(size_t)r * 16 + i] : 0u;
    uint32_t* out = a.out + r;
    uint32_t col = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) out[(size_t)(col++) * H] = s[i];
    external_linear<F>(s);
    auto full_round = [&](int rr) {
        uint32_t reg[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
            uint32_t x = fadd<F>(s[i], k.ext_rc[16 * rr + i]);
            uint32_t x3 = fmul<F>(fmul<F>(x, x), x);
            reg[i] = x3;
            s[i] = R ? fmul<F>(fmul<F>(x3, x3), x) : x3;
        }
        if (R) {
#pragma unroll
            for (int i = 0; i < 16; i++) out[(size_t)(col++) * H] = reg[i];
        }
        external_linear<F>(s);
#pragma unroll
        for (int i = 0; i < 16; i++) out[(size_t)(col++) * H] = s[i];
    };
#pragma unroll 1
    for (int rr = 0; rr < 4; rr++) full_round(rr);
#pragma unroll 1
    for (int rr = 0; rr < F::ROUNDS_P; rr++) {
        uint32_t x = fadd<F>(s[0], k.int_rc[rr]);
        uint32_t x3 = fmul<F>(fmul<F>(x, x), x);
        uint32_t y = R ? fmul<F>(fmul<F>(x3, x3), x) : x3;
        if (R) out[(size_t)(col++) * H] = x3;
        out[(size_t)(col++) * H] = y;
        s[0] = y;
        uint32_t sum = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) sum = fadd<F>(sum, s[i]);
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = fadd<F>(sum, fmul<F>(k.diag[i], s[i]));
    }
#pragma unroll 1
    for (int rr = 4; rr < 8; rr++) full_round(rr);
    // circuit columns
    uint32_t bit = 0, acc = 0;
    if (r < a.n_ops) {
        bit = a.mmcs_bit[r] ? F::R : 0u;
        uint32_t t = r;
        while (t > 0 && a.merkle_path[t] != 0 && a.new_start[t] == 0) t--;
        acc = a.idx_sum[t];
        for (uint32_t u = t + 1; u <= r; u++) acc = fadd<F>(fadd<F>(acc, acc), a.mmcs_bit[u] ? F::R : 0u);
    }
    out[(size_t)(col++) * H] = bit;
    out[(size_t)(col++) * H] = acc;
}

// ------------------------------------------------------------------------------------------------
// Runner assist: the Poseidon2 permutation rows of `CircuitRunner::execute_all`
// (/root/reference circuit/src/tables/runner.rs:256-308 -> circuit/src/ops/poseidon_perm/executor.rs:924-975) executed as CHAINS on
// the device. The host runner does, per operation in order: state = zeros (new_start) or the previous output (sponge mode: the
// whole state; arity-2 Merkle mode: the first RATE_EXT limbs) :111-139, sibling limbs into [RATE_EXT, WIDTH_EXT) on Merkle rows
// :171-208, CTL-exposed witness limbs overwrite :214-225, rate halves swapped when the direction bit is set :233-240, permute.
// A chain (new_start row .. next new_start row) is sequential, chains are independent: one thread per chain. D = 4, width 16
// (WIDTH_EXT 4, RATE_EXT 2). "Previous output" is the previous ROW's output — what the AIR's chaining constraints bind
// (poseidon2-circuit-air/src/air.rs:1024-1081); the executor keeps one slot per mode, which is the same thing whenever the
// rows of a chain are adjacent. Writes the resolved input state and the output state of every row.
// ------------------------------------------------------------------------------------------------
struct P2ChainArgs {
    const uint8_t* new_start;
    const uint8_t* merkle_path;
    const uint8_t* mmcs_bit;
    const uint8_t* witness_mask;   // bit l: limb l (4 words) is a CTL-exposed witness taken from `values`
    const uint32_t* values;        // n_rows x 16: witness limbs where masked; on Merkle rows words 8..15 = the sibling digest
    uint32_t n_rows;
    uint32_t* inputs_out;          // n_rows x 16
    uint32_t* outputs_out;         // n_rows x 16
};
template <class F>
__global__ void __launch_bounds__(64) k_poseidon2_chains(P2ChainArgs a) {
    const uint32_t r0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (r0 >= a.n_rows || !(a.new_start[r0] || r0 == 0)) return;
    uint32_t prev[16];
#pragma unroll
    for (int i = 0; i < 16; i++) prev[i] = 0;
    for (uint32_t r = r0; r < a.n_rows && (r == r0 || !a.new_start[r]); r++) {
        const bool fresh = a.new_start[r] != 0, merkle = a.merkle_path[r] != 0, bit = a.mmcs_bit[r] != 0;
        const uint32_t mask = a.witness_mask[r];
        const uint32_t* v = a.values + (size_t)r * 16;
        uint32_t st[16];
#pragma unroll
        for (int i = 0; i < 16; i++) st[i] = fresh ? 0u : ((merkle && i >= 8) ? 0u : prev[i]);
        if (merkle) {
#pragma unroll
            for (int i = 8; i < 16; i++) st[i] = v[i];
        }
#pragma unroll
        for (int i = 0; i < 16; i++)
            if ((mask >> (i >> 2)) & 1u) st[i] = v[i];
        if (merkle && bit) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t t = st[i];
                st[i] = st[8 + i];
                st[8 + i] = t;
            }
        }
#pragma unroll
        for (int i = 0; i < 16; i++) a.inputs_out[(size_t)r * 16 + i] = st[i];
        poseidon2_permute<F>(st);
#pragma unroll
        for (int i = 0; i < 16; i++) {
            a.outputs_out[(size_t)r * 16 + i] = st[i];
            prev[i] = st[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ALU table fill (AluAir::trace_to_matrix, /root/reference circuit-prover/src/air/alu_air.rs:497-608; layout :22-58,
// columns air/alu_columns.rs:8-46). One thread per table row. Slot (row, lane) holds nothing, one operation [a, b, c, out], or
// (lane 0 only) a packed Horner run of k operations: a, b, c of the first, out of the last, plus the intermediate
// accumulators, the (a_t, c_t) operands of steps 1..k-1 and b^2. The reference's sequential `prev_lane0_out` carry is the
// out value of the previous row's lane-0 slot, so rows are independent. The matrix is zeroed before the launch.
// ------------------------------------------------------------------------------------------------
struct AluFillArgs {
    const uint32_t* kind;
    const uint32_t* first;
    const uint32_t* values;   // n_ops x 4 x 4 Montgomery words
    uint32_t n_slots, lanes, k_max, log_h, wnr;
    uint32_t* out;            // column-major main trace, height H
};
template <class F>
__global__ void __launch_bounds__(128) k_alu_table_fill(AluFillArgs a) {
    constexpr uint32_t D = 4;
    const uint32_t H = 1u << a.log_h;
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= H) return;
    uint32_t* out = a.out + row;
    auto put = [&](uint32_t col, uint32_t v) { out[(size_t)col * H] = v; };
    auto val = [&](uint32_t op, uint32_t which) {
        const uint32_t* p = a.values + ((size_t)op * 4 + which) * D;
        return Ext4{{p[0], p[1], p[2], p[3]}};
    };
    for (uint32_t lane = 0; lane < a.lanes; lane++) {
        const uint32_t slot = row * a.lanes + lane;
        if (slot >= a.n_slots) break;
        const uint32_t k = a.kind[slot];
        if (k == 0) continue;
        const uint32_t first = a.first[slot], last = first + k - 1, m = lane * 4 * D;
#pragma unroll
        for (uint32_t w = 0; w < 3; w++) {
            const Ext4 v = val(first, w);
#pragma unroll
            for (uint32_t c = 0; c < D; c++) put(m + w * D + c, v.c[c]);
        }
        const Ext4 o = val(last, 3);
#pragma unroll
        for (uint32_t c = 0; c < D; c++) put(m + 3 * D + c, o.c[c]);
        if (k >= 2 && lane == 0) {
            const uint32_t extra = a.lanes * 4 * D, num_int = (a.k_max - 1) / 2;
            const uint32_t ac_base = extra + num_int * D, bsq_base = ac_base + 2 * (a.k_max - 1) * D;
            // accumulator entering this row = out of the previous row's lane-0 slot (0 after a separator / at row 0)
            Ext4 acc = ext_zero();
            if (row > 0) {
                const uint32_t ps = (row - 1) * a.lanes, pk = a.kind[ps];
                if (pk) acc = val(a.first[ps] + pk - 1, 3);
            }
            const Ext4 b = val(first, 1);
            uint32_t step = 0;
            for (uint32_t si = 0; si < num_int; si++) {
                const uint32_t i0 = first + step, i1 = i0 + 1;
                // acc <- acc*b + c - a  (HornerAcc), twice when the second operation belongs to the run
                acc = esub<F>(eadd<F>(emul<F>(acc, b, a.wnr), val(i0, 2)), val(i0, 0));
                step++;
                if (i1 < first + k) {
                    acc = esub<F>(eadd<F>(emul<F>(acc, b, a.wnr), val(i1, 2)), val(i1, 0));
                    step++;
                }
#pragma unroll
                for (uint32_t c = 0; c < D; c++) put(extra + si * D + c, acc.c[c]);
            }
            for (uint32_t t = 1; t < k; t++) {
                const Ext4 at = val(first + t, 0), ct = val(first + t, 2);
#pragma unroll
                for (uint32_t c = 0; c < D; c++) {
                    put(ac_base + 2 * (t - 1) * D + c, at.c[c]);
                    put(ac_base + (2 * (t - 1) + 1) * D + c, ct.c[c]);
                }
            }
            const Ext4 bb = emul<F>(b, b, a.wnr);
#pragma unroll
            for (uint32_t c = 0; c < D; c++) put(bsq_base + c, bb.c[c]);
        }
    }
}

// Synthetic data for the isolated commit benchmark: splitmix64(seed + index) reduced mod P.
template <class F>
__global__ void k_fill_random(uint32_t* out, size_t n, uint64_t seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    out[i] = (uint32_t)(z % F::P);
}

}  // namespace p3r
