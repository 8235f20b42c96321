// p3r.cu — host side of libp3r_b200.so: context, device arena, LDE planner, MMCS commit, the phase-stepped proving
// session behind include/p3r.h, the host DuplexChallenger and the one-shot p3r_prove (host mirror of
// BatchStarkProver::prove, /root/reference circuit-prover/src/batch_stark_prover.rs:1275-1642 -> p3_batch_stark::prove_batch).
// No CPU fallback: every compute entry point runs CUDA kernels from kernels.cuh or fails.
#include <algorithm>
#include <sched.h>
#include <sys/prctl.h>
#include <time.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "ntt_col.cuh"
#include "spec.h"

using namespace p3r;

#define CUDA_TRY(x)                                                                              \
    do {                                                                                         \
        cudaError_t e_ = (x);                                                                    \
        if (e_ != cudaSuccess) {                                                                 \
            set_err(ctx, std::string(#x) + ": " + cudaGetErrorString(e_));                       \
            return P3R_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)
#define TRY(x)                   \
    do {                         \
        int rc_ = (x);           \
        if (rc_ != P3R_OK) return rc_; \
    } while (0)

// ------------------------------------------------------------------------------------------------
struct Slab {
    char* base = nullptr;
    size_t size = 0, used = 0;
};
// Host wait for a stream. Mode 1 (default): poll cudaStreamQuery + sched_yield, with every device->host result staged through
// pinned memory (d2h_async / ctx_wait) so that no call blocks inside the driver: a waiting thread gives its core to a thread
// that has kernels to launch. Mode 0: cudaStreamSynchronize and direct copies (the driver spins; with a pageable destination
// the spin happens inside cudaMemcpyAsync). Mode 2: block on a cudaEventBlockingSync event (the thread sleeps; 0.8 ms of CPU
// per proof instead of 3.9, but each wake-up costs latency). Measured on one B200, four proofs in flight: yield 412 proofs/s
// (406 with the process confined to 2 cores), spin 379-397 (314), block 350 (351). Process-wide; P3R_WAIT=spin|yield|block or
// p3r_set_wait_mode().
static std::atomic<int> g_wait_mode{-1};
static inline int wait_mode() {
    int m = g_wait_mode.load(std::memory_order_relaxed);
    if (m >= 0) return m;
    const char* e = getenv("P3R_WAIT");
    m = !e ? 1 : (!strcmp(e, "block") ? 2 : (!strcmp(e, "spin") ? 0 : (!strcmp(e, "sleep") ? 3 : 1)));
    g_wait_mode.store(m, std::memory_order_relaxed);
    return m;
}
static inline cudaError_t stream_wait(cudaStream_t s) {
    const int m = wait_mode();
    if (m == 0) return cudaStreamSynchronize(s);
    if (m == 1) {
        for (;;) {
            cudaError_t e = cudaStreamQuery(s);
            if (e != cudaErrorNotReady) return e;
            sched_yield();
        }
    }
    if (m == 3) {
        // Poll, but give the core away for real between polls: a short yielding phase (results that are about to arrive), then
        // 30 us sleeps (timer slack lowered to 1 us for this thread). For hosts with fewer cores than proving threads — 8 ranks x 4
        // lanes on 32 cores: with the pure yield loop every core runs a poller and the threads that have kernels to launch or
        // proofs to hand over queue behind them (aggregation tree at 8 GPUs: 0.81 of linear with yield).
        thread_local bool slack_set = false;
        if (!slack_set) {
            prctl(PR_SET_TIMERSLACK, 1000UL, 0, 0, 0);
            slack_set = true;
        }
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            cudaError_t e = cudaStreamQuery(s);
            if (e != cudaErrorNotReady) return e;
            if (std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(20)) {
                sched_yield();
            } else {
                struct timespec ts = {0, 30000};
                nanosleep(&ts, nullptr);
            }
        }
    }
    thread_local cudaEvent_t ev = nullptr;   // one device per process (one rank per GPU), so one event per host thread
    if (!ev) {
        cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaEventRecord(ev, s);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ev);
}

struct Arena {  // bump allocator over cudaMalloc'd slabs; reset() keeps the slabs for the next session
    std::vector<Slab> slabs;
    size_t cur = 0;
    void reset() {
        for (auto& s : slabs) s.used = 0;
        cur = 0;
    }
    void* alloc(size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        for (; cur < slabs.size(); cur++) {
            Slab& s = slabs[cur];
            if (s.used + bytes <= s.size) {
                void* p = s.base + s.used;
                s.used += bytes;
                return p;
            }
        }
        Slab s;
        s.size = std::max(bytes, (size_t)256 << 20);
        if (cudaMalloc(&s.base, s.size) != cudaSuccess) return nullptr;
        s.used = bytes;
        slabs.push_back(s);
        cur = slabs.size() - 1;
        return s.base;
    }
    void destroy() {
        for (auto& s : slabs) cudaFree(s.base);
        slabs.clear();
    }
};
struct GTable {
    uint32_t *lo = nullptr, *hi = nullptr;
};

enum KClass { KC_NTT = 0, KC_HASH, KC_COMPRESS, KC_LOGUP, KC_QUOTIENT, KC_OPEN, KC_REDUCE, KC_FOLD, KC_TRANSPOSE, KC_MISC, KC_COUNT };
static const char* const KCLASS_NAMES[KC_COUNT] = {"ntt_lde", "hash_rows", "compress", "logup", "quotient", "open", "reduced_openings",
                                                   "fri_fold", "transpose", "misc"};
struct KernelStats {
    double ms[KC_COUNT] = {0};
    uint64_t launches[KC_COUNT] = {0};
    uint64_t bytes[KC_COUNT] = {0};  // algorithmic bytes (DESIGN.md "Kernels")
    uint64_t perms[KC_COUNT] = {0};  // Poseidon2 permutations issued (hash_rows / compress classes)
};

struct p3r_ctx {
    bool use_spec = true;  // p3r_set_specialization
    cudaEvent_t timer_ev[2] = {nullptr, nullptr};
    uint32_t time_mask = 0;  // bit per KClass: record CUDA events around launches of that class
    std::vector<cudaEvent_t> ev_pool;
    std::vector<cudaEvent_t> phase_ev;  // PhaseTimer's events
    size_t ev_used = 0;
    std::vector<std::pair<int, size_t>> ev_pending;  // (class, index of start event)
    KernelStats kstats;
    int device = 0;
    int field_id = 0;
    p3r_field_desc field{};
    p3r_fri_params fri{};
    Poseidon2Consts p2{};
    uint32_t w_m = 0, gen_m = 0, inv2_m = 0;  // Montgomery
    cudaStream_t stream = nullptr;
    // Side streams for independent kernels of one phase (the per-table quotient / LogUp kernels are latency-bound and touch
    // disjoint data): fork_streams() makes them wait for everything queued on `stream`, join_streams() the reverse.
    static constexpr int N_AUX = 4;
    cudaStream_t aux[N_AUX] = {nullptr, nullptr, nullptr, nullptr};
    int stream_prio = 0;      // CUDA priority of `stream` and the side streams (p3r_ctx_set_stream_priority)
    cudaEvent_t aux_ev[N_AUX + 1] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    uint32_t* tws = nullptr;  // per-stage compact twiddle tables, 2^logT - 1 entries
    uint32_t* tw = nullptr;   // = tws + 2^(logT-1) - 1: half table of omega_T
    void (*host_permute)(uint32_t*, const Poseidon2Consts&) = nullptr;  // transcript permutation (AVX2 or scalar), set at creation
    bool dev_fri_transcript = true;  // FRI commit rounds without host round trips (p3r_set_specialization bit 2 turns it off)
    bool use_grouped_interp = true; // k_quotient_grouped for long programs without a generated kernel (p3r_set_specialization bit 4 off)
    uint32_t lde_streams = 2;       // job groups (streams) of one batched LDE, 1..N_AUX (P3R_LDE_STREAMS); measured best: 2
    bool lde_small_cta = false;     // 2^14-element CTAs (two per SM) for columns of up to 2^14 rows: P3R_LDE_SMALL_CTA=1; measured
                                    // neutral (LDE class 0.398 vs 0.392 ms per layer proof), so the single CTA shape stays the default
    bool uni_stark = false;         // p3r_ctx_set_uni_stark: single-table proofs with p3-uni-stark's transcript head
    p3r_conventions conv{0, 0, 0};  // p3r_ctx_set_conventions
    bool use_hash_queue = false;  // work-queue row hashing (p3r_set_specialization bit 3 turns it ON; measured slower, see kernels.cuh)
    uint32_t n_sms = 148;
    bool use_col_ntt = true;  // whole-column LDE kernels for 2^5..2^15 rows (p3r_set_specialization bit 1 turns them off)
    uint32_t logT = 0;
    uint32_t r4 = 0, r8 = 0, r8_3 = 0;
    Poseidon2ConstsW* d_p2w = nullptr;  // width-24 leaf hasher (p3r_ctx_set_leaf_hasher), nullptr = width-16 sponge
    Poseidon2Consts* d_p2 = nullptr;  // global-memory copy of the Poseidon2 constants (per-lane reads of the cooperative kernels)
    std::map<uint32_t, GTable> gtables;
    Arena arena;
    // pinned staging for small uploads/downloads
    char* pin = nullptr;
    size_t pin_size = 0, pin_used = 0, pin_top_used = 0;
    size_t mirror_valid = 0;        // dstage[0, mirror_valid) == pin[0, mirror_valid): see upload_small
    bool skip_equal_uploads = true; // P3R_UPLOAD_SKIP=0 turns the skipping off (A/B)
    uint64_t uploads_skipped = 0;
    // D2H staging for the yield / block wait modes (d2h_async, ctx_wait)
    char* pin_out = nullptr;
    size_t pin_out_size = 0, pin_out_used = 0;
    struct PendingCopy {
        void* dst;
        const void* src;
        size_t bytes;
    };
    std::vector<PendingCopy> pending;
    uint32_t *grind_pin = nullptr, *grind_dev = nullptr;  // private 128-byte staging of p3r_grind
    char* dstage = nullptr;  // device mirror for small uploads
    size_t dstage_size = 0;
    std::string err;
    uint64_t launches = 0;
    std::vector<std::pair<std::string, float>> phase_times;
    std::string phase_names_blob;
};
static void set_err(p3r_ctx* ctx, const std::string& s) {
    if (ctx) ctx->err = s;
}

#define LAUNCH_CHECK_C(cls)                                              \
    do {                                                                 \
        ctx->kstats.launches[cls]++;                                     \
        LAUNCH_CHECK();                                                  \
    } while (0)
#define LAUNCH_CHECK()                                                   \
    do {                                                                 \
        ctx->launches++;                                                 \
        cudaError_t e_ = cudaGetLastError();                             \
        if (e_ != cudaSuccess) {                                         \
            set_err(ctx, std::string("kernel launch: ") + cudaGetErrorString(e_)); \
            return P3R_ERR_CUDA;                                         \
        }                                                                \
    } while (0)

// Scoped CUDA-event pair around the launches issued while it is alive (only when the class is enabled in time_mask).
struct KT {
    p3r_ctx* ctx;
    int cls;
    bool on;
    KT(p3r_ctx* c, int k, uint64_t algo_bytes = 0) : ctx(c), cls(k) {
        ctx->kstats.bytes[k] += algo_bytes;
        on = (ctx->time_mask >> k) & 1u;
        if (!on) return;
        if (ctx->ev_used + 2 > ctx->ev_pool.size()) {
            for (int i = 0; i < 256; i++) {
                cudaEvent_t e;
                cudaEventCreate(&e);
                ctx->ev_pool.push_back(e);
            }
        }
        ctx->ev_pending.push_back({k, ctx->ev_used});
        cudaEventRecord(ctx->ev_pool[ctx->ev_used], ctx->stream);
        ctx->ev_used += 2;
    }
    ~KT() {
        if (on) cudaEventRecord(ctx->ev_pool[ctx->ev_pending.back().second + 1], ctx->stream);
    }
};
static cudaError_t ctx_wait(p3r_ctx* ctx);
static void kstats_collect(p3r_ctx* ctx) {
    if (ctx->ev_pending.empty()) return;
    ctx_wait(ctx);
    for (auto& pr : ctx->ev_pending) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev_pool[pr.second], ctx->ev_pool[pr.second + 1]);
        ctx->kstats.ms[pr.first] += ms;
    }
    ctx->ev_pending.clear();
    ctx->ev_used = 0;
}

template <class T>
static T* arena_alloc(p3r_ctx* ctx, size_t count) {
    return reinterpret_cast<T*>(ctx->arena.alloc(count * sizeof(T)));
}
// Device -> host copy of a result the host reads after the next ctx_wait(). With the driver's spin wait (mode 0) it is a plain
// cudaMemcpyAsync: for a pageable destination the driver itself spins until the data has arrived. In the yield / block modes
// the copy lands in pinned staging (truly asynchronous) and ctx_wait() moves it to `dst` after the stream has drained, so the
// thread really sleeps while the GPU works.
static cudaError_t d2h_async(p3r_ctx* ctx, void* dst, const void* dsrc, size_t bytes) {
    const size_t b = (bytes + 63) & ~(size_t)63;
    if (wait_mode() == 0 || !ctx->pin_out || ctx->pin_out_used + b > ctx->pin_out_size)
        return cudaMemcpyAsync(dst, dsrc, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    char* h = ctx->pin_out + ctx->pin_out_used;
    ctx->pin_out_used += b;
    ctx->pending.push_back({dst, h, bytes});
    return cudaMemcpyAsync(h, dsrc, bytes, cudaMemcpyDeviceToHost, ctx->stream);
}
static cudaError_t ctx_wait(p3r_ctx* ctx) {
    cudaError_t e = stream_wait(ctx->stream);
    for (auto& c : ctx->pending) std::memcpy(c.dst, c.src, c.bytes);
    ctx->pending.clear();
    ctx->pin_out_used = 0;
    return e;
}
// Copy a small host blob to the device through the pinned staging area (valid until the next session begins).
// The ring is reset at the start of every session and a circuit shape allocates it in the same order every time, so a slot
// usually receives the bytes it already holds (job descriptors, column pointers, twiddle tables: everything that does not depend
// on the proof's challenges). The pinned copy of the previous session doubles as the record of what the device mirror holds:
// inside the prefix [0, mirror_valid) the mirror is known to equal the pinned ring, and when the new bytes equal the pinned ones
// the host-to-device copy is skipped (most of the ~40 small copies of a layer proof, each a stream operation of its own between
// two kernels). `device_mutable`: a kernel writes to the device copy (challenger state, queue counters); those slots come from
// the top of the ring, are always copied and never trusted.
static void ring_reset(p3r_ctx* ctx) {
    ctx->pin_used = 0;
    ctx->pin_top_used = 0;
}
static void* upload_small(p3r_ctx* ctx, const void* src, size_t bytes, bool device_mutable = false) {
    const size_t b = (bytes + 255) & ~(size_t)255;
    if (ctx->pin_used + ctx->pin_top_used + b > ctx->pin_size) return nullptr;
    size_t off;
    if (device_mutable) {
        ctx->pin_top_used += b;
        off = ctx->pin_size - ctx->pin_top_used;
    } else {
        off = ctx->pin_used;
        ctx->pin_used += b;
        if (ctx->skip_equal_uploads && off + b <= ctx->mirror_valid && std::memcmp(ctx->pin + off, src, bytes) == 0) {
            ctx->uploads_skipped++;
            return ctx->dstage + off;
        }
    }
    char* h = ctx->pin + off;
    char* d = ctx->dstage + off;
    std::memcpy(h, src, bytes);
    if (cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return nullptr;
    if (!device_mutable && off <= ctx->mirror_valid) ctx->mirror_valid = std::max(ctx->mirror_valid, off + b);
    if (device_mutable) ctx->mirror_valid = std::min(ctx->mirror_valid, off);   // (only if the two ends ever met)
    return d;
}
template <class T>
static T* upload_vec(p3r_ctx* ctx, const std::vector<T>& v) {
    if (v.empty()) return reinterpret_cast<T*>(ctx->dstage);
    return reinterpret_cast<T*>(upload_small(ctx, v.data(), v.size() * sizeof(T)));
}

static int fork_streams(p3r_ctx* ctx) {
    for (int i = 0; i < p3r_ctx::N_AUX; i++)
        if (!ctx->aux[i]) {
            CUDA_TRY(cudaStreamCreateWithPriority(&ctx->aux[i], cudaStreamNonBlocking, ctx->stream_prio));
            CUDA_TRY(cudaEventCreateWithFlags(&ctx->aux_ev[i], cudaEventDisableTiming));
        }
    if (!ctx->aux_ev[p3r_ctx::N_AUX]) CUDA_TRY(cudaEventCreateWithFlags(&ctx->aux_ev[p3r_ctx::N_AUX], cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(ctx->aux_ev[p3r_ctx::N_AUX], ctx->stream));
    for (int i = 0; i < p3r_ctx::N_AUX; i++) CUDA_TRY(cudaStreamWaitEvent(ctx->aux[i], ctx->aux_ev[p3r_ctx::N_AUX], 0));
    return P3R_OK;
}
static int join_streams(p3r_ctx* ctx) {
    for (int i = 0; i < p3r_ctx::N_AUX; i++) {
        CUDA_TRY(cudaEventRecord(ctx->aux_ev[i], ctx->aux[i]));
        CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->aux_ev[i], 0));
    }
    return P3R_OK;
}

// ------------------------------------------------------------------------------------------------
// device-side descriptions
// ------------------------------------------------------------------------------------------------
struct MatRef {  // column-major device matrix
    const uint32_t* d = nullptr;
    uint32_t log_h = 0, w = 0;
};
struct Tree {
    uint32_t log_max_h = 0;
    uint32_t* digests = nullptr;  // layers back to back, level l at digest offset 2^(log_max_h+1) - 2^(log_max_h-l+1)
    std::vector<MatRef> mats;     // commit order
    size_t level_off(uint32_t l) const { return ((size_t)2 << log_max_h) - ((size_t)2 << (log_max_h - l)); }
};
struct InstDev {
    uint32_t log_h = 0, main_w = 0, prep_w = 0, n_pub = 0, log_qc = 0, uses_next = 0;
    uint4* cons = nullptr;
    uint32_t n_cons_insns = 0, n_constraints = 0;
    Ext4* cons_econst = nullptr;
    uint4* lk = nullptr;
    uint32_t n_lk_insns = 0, n_lk_outs = 0;
    std::vector<p3r_lookup> lookups;
    std::vector<p3r_interaction> inter;
    p3r_lookup* d_lookups = nullptr;
    p3r_interaction* d_inter = nullptr;
    uint32_t* prep_trace = nullptr;  // natural order, column-major
    uint32_t* prep_lde = nullptr;
    uint32_t* sel = nullptr;
    uint32_t* inv_van = nullptr;
    uint4* gcons = nullptr;             // the constraint program cut into QG_GROUPS sub-programs (programs of >= 256 instructions)
    uint32_t goff[QG_GROUPS + 1] = {0};
    SpecQuotientKernel spec = nullptr;  // build-time specialised quotient kernel whose program hash matches, if any
    SpecLogupKernel spec_logup = nullptr;  // same for the LogUp trace rows
    uint32_t aux_w() const { return lookups.empty() ? 0 : (uint32_t)lookups.size() + 1; }
};
struct p3r_prep {
    p3r_ctx* ctx = nullptr;
    std::vector<InstDev> inst;
    std::vector<void*> owned;  // cudaMalloc'd
    Tree prep_tree;
    bool has_prep = false, has_perm = false;
    uint32_t max_msg_w = 1, n_buses = 0;
    std::vector<uint32_t> prep_cap;  // host copy, Montgomery
};

struct p3r_traces {  // device-resident column-major copies of one layer's traces
    p3r_ctx* ctx = nullptr;
    std::vector<uint32_t*> d;   // per instance, pointers into `slab`
    void* slab = nullptr;       // one allocation for all instances
};

enum Phase { PH_BEGIN = 0, PH_MAIN, PH_PERM, PH_QUOT, PH_OPEN, PH_FRI };

struct FriRound {
    uint32_t log_arity = 0, log_len = 0;  // length of the vector committed in this round
    Ext4* vec = nullptr;                  // committed vector (bit-reversed EF, AoS) = row-major matrix arity*4 wide
    Tree tree;
};
struct p3r_session {
    p3r_ctx* ctx = nullptr;
    const p3r_prep* prep = nullptr;
    int phase = PH_BEGIN;
    std::vector<uint32_t*> trace, main_lde, perm, perm_lde, chunks, chunk_lde, d_pub;
    uint32_t* scratch_coef = nullptr;   // LDE coefficient scratch
    uint32_t* scratch_tmp = nullptr;    // LDE multi-pass scratch
    Tree main_tree, perm_tree, quot_tree;
    Ext4* d_chal = nullptr;             // per instance block of [prefix, beta] pairs
    std::vector<uint32_t> chal_off;
    Ext4* d_terminals = nullptr;        // one per instance
    Ext4* d_opened = nullptr;
    size_t n_opened = 0;                // Ext4 count
    std::vector<uint32_t> opened_host;  // Montgomery words
    Ext4 zeta{};
    // FRI
    std::vector<uint32_t> heights;      // distinct LDE log heights, descending
    std::map<uint32_t, Ext4*> ro;
    std::vector<FriRound> rounds;
    Ext4* final_vec = nullptr;
    uint32_t log_max = 0;
    // openings bookkeeping: for every (round, matrix) the offsets of its opened values in d_opened
    struct OpenRef {
        uint32_t inst, kind;  // kind 0 main, 1 quotient chunk, 2 prep, 3 perm
        uint32_t chunk;
        uint32_t off[2];
        uint32_t n_points;
        uint32_t width;
    };
    std::vector<std::vector<OpenRef>> open_rounds;  // [main, quot, prep?, perm?]
};

// ------------------------------------------------------------------------------------------------
// Host DuplexChallenger — product code, Montgomery arithmetic, shares nothing with oracle/.
// ------------------------------------------------------------------------------------------------
// host_p2_avx2.cpp (the only translation unit built with -mavx2)
namespace p3r {
void host_permute_avx2_koalabear(uint32_t* st, const Poseidon2Consts& k);
void host_permute_avx2_babybear(uint32_t* st, const Poseidon2Consts& k);
}  // namespace p3r
template <class F>
static void host_permute_scalar(uint32_t* st, const Poseidon2Consts& k) {
    poseidon2_permute_with<F>(st, k);
}
typedef void (*HostPermuteFn)(uint32_t*, const Poseidon2Consts&);
// AVX2 permutation when the CPU has it and it reproduces the scalar twin on a probe state; the scalar twin otherwise.
template <class F>
static HostPermuteFn pick_host_permute(const Poseidon2Consts& k) {
    HostPermuteFn scalar = host_permute_scalar<F>;
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx2")) {
        HostPermuteFn fast = FieldId<F>::value == 0 ? host_permute_avx2_koalabear : host_permute_avx2_babybear;
        uint32_t a[16], b[16];
        for (int i = 0; i < 16; i++) a[i] = b[i] = (uint32_t)((i + 1) * 0x9E3779B1u) % F::P;
        for (int it = 0; it < 3; it++) {
            fast(a, k);
            scalar(b, k);
        }
        if (std::memcmp(a, b, sizeof a) == 0) return fast;
    }
#endif
    return scalar;
}

template <class F>
struct HostChallenger {
    const Poseidon2Consts* k;
    HostPermuteFn permute;
    uint32_t st[16];
    uint32_t in[8], out[8];
    int n_in = 0, n_out = 0;
    double host_ms = 0;  // wall time spent in host permutations
    uint32_t n_perms = 0;
    explicit HostChallenger(const Poseidon2Consts* k_, HostPermuteFn fn = nullptr)
        : k(k_), permute(fn ? fn : host_permute_scalar<F>) {
        std::memset(st, 0, sizeof st);
    }
    void duplex() {
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < n_in; i++) st[i] = in[i];
        if (n_in > 0) {
            for (int i = n_in; i < 8; i++) st[i] = 0;
            st[8] = fadd<F>(st[8], to_monty<F>((uint32_t)n_in));
        }
        n_in = 0;
        permute(st, *k);
        for (int i = 0; i < 8; i++) out[i] = st[i];
        n_out = 8;
        n_perms++;
        host_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    void observe(uint32_t v) {
        n_out = 0;
        in[n_in++] = v;
        if (n_in == 8) duplex();
    }
    void observe_words(const uint32_t* v, size_t n) {
        for (size_t i = 0; i < n; i++) observe(v[i]);
    }
    void observe_lifted(uint32_t canonical) {  // observe_base_as_algebra_element
        observe(to_monty<F>(canonical));
        observe(0);
        observe(0);
        observe(0);
    }
    uint32_t sample() {
        if (n_in > 0 || n_out == 0) duplex();
        return out[--n_out];
    }
    void sample_ext(uint32_t e[4]) {
        for (int i = 0; i < 4; i++) e[i] = sample();
    }
    uint32_t sample_bits(uint32_t bits) { return from_monty<F>(sample()) & ((1u << bits) - 1); }
};

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
template <class F>
static int ensure_twiddles(p3r_ctx* ctx, uint32_t logT) {
    if (logT < 12) logT = 12;
    if (ctx->tw && ctx->logT >= logT) return P3R_OK;
    if (logT > F::TWO_ADICITY) {
        set_err(ctx, "domain exceeds the field's two-adicity");
        return P3R_ERR_INVALID_ARG;
    }
    CUDA_TRY(ctx_wait(ctx));
    if (ctx->tws) cudaFree(ctx->tws);
    ctx->tws = ctx->tw = nullptr;
    size_t count = ((size_t)1 << logT) - 1;
    CUDA_TRY(cudaMalloc(&ctx->tws, (count + 1) * 4));
    uint32_t gen = ctx->gen_m;
    uint32_t w = fpow<F>(gen, ((uint64_t)F::P - 1) >> logT);
    k_stage_twiddles<F><<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(ctx->tws, logT, w);
    LAUNCH_CHECK();
    ctx->tw = ctx->tws + (((size_t)1 << (logT - 1)) - 1);
    ctx->logT = logT;
    ctx->r4 = fpow<F>(w, (uint64_t)1 << (logT - 2));
    ctx->r8 = fpow<F>(w, (uint64_t)1 << (logT - 3));
    ctx->r8_3 = fpow<F>(w, (uint64_t)3 << (logT - 3));
    for (auto& kv : ctx->gtables) {  // tables do not depend on logT, keep them
        (void)kv;
    }
    return P3R_OK;
}
template <class F>
static int get_gtable(p3r_ctx* ctx, uint32_t log_n, GTable* out) {
    auto it = ctx->gtables.find(log_n);
    if (it != ctx->gtables.end()) {
        *out = it->second;
        return P3R_OK;
    }
    GTable t;
    uint32_t n_hi = log_n > 10 ? (1u << (log_n - 10)) : 1;
    CUDA_TRY(cudaMalloc(&t.lo, 1024 * 4));
    CUDA_TRY(cudaMalloc(&t.hi, n_hi * 4));
    uint32_t n_inv = finv<F>(to_monty<F>(1u << log_n));
    uint32_t g = ctx->gen_m;
    k_powers<F><<<4, 256, 0, ctx->stream>>>(t.lo, 1024, g, n_inv);
    LAUNCH_CHECK();
    k_powers<F><<<(n_hi + 255) / 256, 256, 0, ctx->stream>>>(t.hi, n_hi, fpow<F>(g, 1024), F::R);
    LAUNCH_CHECK();
    ctx->gtables[log_n] = t;
    *out = t;
    return P3R_OK;
}

// ------------------------------------------------------------------------------------------------
// K4 host planner: batched coset LDE of `w` columns (natural order, column-major, height n) into dst
// (column-major, height N = n << log_blowup, rows bit-reversed). in_shift = GENERATOR^(1-use_g) * w_N^{-rot}... see NttPass.
// scratch_coef: n*w words; scratch_tmp: N*w words (only touched when log_n > TILE_LOG).
// ------------------------------------------------------------------------------------------------
constexpr uint32_t TILE_LOG = 13;
// Merkle levels up to this many nodes go through the fused k_merkle_stage launches (P3R_STAGE_MAX_LOG overrides the log2)
static const uint32_t STAGE_MAX_NODES = [] {
    const char* e = getenv("P3R_STAGE_MAX_LOG");
    uint32_t l = e ? (uint32_t)atoi(e) : 12u;   // measured: compress class 1.03 ms at 2^12 against 1.06 at 2^13, 1.17 at 2^14
    return 1u << std::max(8u, std::min(l, 16u));
}();
struct PassPlan {
    uint32_t s0, r, log_cw;
};
static std::vector<PassPlan> plan_passes(uint32_t log_n) {
    std::vector<PassPlan> p;
    if (log_n <= TILE_LOG) {
        p.push_back({0, log_n, 0});
        return p;
    }
    // contiguous first chunk, then strided chunks of <= 8 stages (tile rows <= 256, >= 32 consecutive elements per row)
    uint32_t n_strided = (log_n - TILE_LOG + 7) / 8;
    uint32_t strided_total = log_n - std::min(log_n, TILE_LOG);
    // balance: last chunk at least 5 stages so bit-reversed stores form >=128-byte runs
    std::vector<uint32_t> chunks;
    uint32_t first = log_n - strided_total;
    if (strided_total < 5) {
        first -= (5 - strided_total);
        strided_total = 5;
        n_strided = 1;
    }
    chunks.push_back(first);
    for (uint32_t i = 0; i < n_strided; i++) {
        uint32_t c = strided_total / (n_strided - i);
        chunks.push_back(c);
        strided_total -= c;
    }
    uint32_t s0 = 0;
    for (size_t i = 0; i < chunks.size(); i++) {
        uint32_t r = chunks[i];
        uint32_t log_cw = (i == 0) ? (TILE_LOG - r) : std::min(s0, TILE_LOG - r);
        if (i == 0) log_cw = std::min(log_cw, log_n - r);
        p.push_back({s0, r, log_cw});
        s0 += r;
    }
    return p;
}
struct LdeJob {
    const uint32_t* src;   // natural order, column-major, height n
    uint32_t* dst;         // column-major, height n << log_blowup, bit-reversed rows
    uint32_t log_n, w;
    bool use_g;
    uint32_t rot;
    uint32_t* coef;        // n*w words of scratch
    uint32_t* tmp;         // (n << log_blowup)*w words of scratch, only when log_n > TILE_LOG
};
template <class F>
static int launch_pass_level(p3r_ctx* ctx, std::vector<NttPass>& passes) {
    if (passes.empty()) return P3R_OK;
    size_t smem = 0;
    uint32_t cta = 0;
    for (auto& a : passes) {
        uint32_t R = 1u << a.r, CW = 1u << a.log_cw;
        size_t total = (size_t)R * CW;
        smem = std::max(smem, (a.s0 == 0 ? total + (total >> 5) + 1 : (size_t)R * (CW + 1)) * 4);
        a.n_tiles = (1u << a.log_n) / (R * CW);
        a.cta_begin = cta;
        cta += a.n_tiles * a.n_cols * a.n_cosets;
    }
    static bool attr_set[2] = {false, false};
    int fid = FieldId<F>::value;
    if (!attr_set[fid]) {
        cudaFuncSetAttribute(k_ntt_pass<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr_set[fid] = true;
    }
    NttPass* d_jobs = upload_vec(ctx, passes);
    if (!d_jobs) {
        set_err(ctx, "staging exhausted");
        return P3R_ERR_OOM;
    }
    k_ntt_pass<F><<<cta, 512, smem, ctx->stream>>>(d_jobs, (uint32_t)passes.size());
    LAUNCH_CHECK_C(KC_NTT);
    return P3R_OK;
}
// Whole-column path (ntt_col.cuh) for jobs with 2^5 <= n <= 2^15: one inverse and one forward launch per size class.
static void col_plan(uint32_t log_n, uint32_t q[3]) {
    static const uint8_t plans[11][3] = {{0, 0, 0}, {1, 0, 0}, {2, 0, 0}, {3, 0, 0}, {4, 0, 0}, {2, 3, 0},
                                         {3, 3, 0}, {3, 4, 0}, {4, 4, 0}, {3, 3, 3}, {3, 3, 4}};
    for (int i = 0; i < 3; i++) q[i] = plans[log_n - 5][i];
}
template <class F>
static int coset_lde_cols(p3r_ctx* ctx, const std::vector<LdeJob>& jobs, uint32_t log_blowup) {
    if (jobs.empty()) return P3R_OK;
    static bool attr_set[2] = {false, false};
    const int fid = FieldId<F>::value;
    if (!attr_set[fid]) {
        cudaFuncSetAttribute(k_ntt_col<F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 << COL_MAX_LOG);
        cudaFuncSetAttribute(k_ntt_col<F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 << COL_MAX_LOG);
        attr_set[fid] = true;
    }
    const uint32_t n_cosets = 1u << log_blowup;
    // roots needed on the host: w_32 powers and w_N per job
    const uint32_t w32 = fpow<F>(ctx->gen_m, ((uint64_t)F::P - 1) >> 5);
    uint32_t st32[31], ist32[31];  // w_{2^(s+1)}^e and its inverse at (2^s - 1) + e, s < 5
    for (uint32_t sgm = 0; sgm < 5; sgm++)
        for (uint32_t e = 0; e < (1u << sgm); e++) {
            uint32_t w = fpow<F>(w32, (uint64_t)e << (4 - sgm));
            st32[(1u << sgm) - 1 + e] = w;
            ist32[(1u << sgm) - 1 + e] = finv<F>(w);
        }
    std::vector<uint32_t> ctab;  // [inverse table (1 entry)] then per job, per coset
    ctab.resize(COL_CTAB, 0);
    for (int i = 0; i < 31; i++) ctab[i] = ist32[i];
    std::vector<size_t> job_ctab(jobs.size());
    for (size_t qi = 0; qi < jobs.size(); qi++) {
        const LdeJob& j = jobs[qi];
        const uint32_t logN = j.log_n + log_blowup;
        const uint32_t wN = fpow<F>(ctx->gen_m, ((uint64_t)F::P - 1) >> logN);
        job_ctab[qi] = ctab.size();
        for (uint32_t cs = 0; cs < n_cosets; cs++) {
            const uint32_t rj = bitrev32(cs, log_blowup);
            const uint64_t e = ((uint64_t)rj + j.rot) & (((uint64_t)1 << logN) - 1);
            uint32_t c = fpow<F>(wN, e);
            if (j.use_g) c = fmul<F>(c, ctx->gen_m);
            uint32_t C[28];
            for (uint32_t sgm = 0; sgm < 28; sgm++) C[sgm] = F::R;
            C[j.log_n - 1] = c;
            for (uint32_t sgm = j.log_n - 1; sgm-- > 0;) C[sgm] = fmul<F>(C[sgm + 1], C[sgm + 1]);
            size_t off = ctab.size();
            ctab.resize(off + COL_CTAB);
            for (uint32_t sgm = 0; sgm < 5; sgm++)
                for (uint32_t ee = 0; ee < (1u << sgm); ee++)
                    ctab[off + (1u << sgm) - 1 + ee] = fmul<F>(C[sgm], st32[(1u << sgm) - 1 + ee]);
            for (uint32_t sgm = 0; sgm < 28; sgm++) ctab[off + 31 + sgm] = C[sgm];
        }
    }
    const uint32_t* d_ctab = upload_vec(ctx, ctab);
    if (!d_ctab) {
        set_err(ctx, "staging exhausted");
        return P3R_ERR_OOM;
    }
    ColJob base{};
    base.tws = ctx->tws;
    base.r4 = ctx->r4;
    base.r8 = ctx->r8;
    base.r8_3 = ctx->r8_3;
    {
        const uint32_t w16 = fmul<F>(w32, w32);
        uint32_t p = F::R;
        for (int e = 0; e < 8; e++) {
            base.r16[e] = p;
            p = fmul<F>(p, w16);
        }
    }
    // one inverse and one forward launch for all jobs: every CTA works on up to 2^15 elements (2^(15 - log_n) columns),
    // largest columns first. Columns of 2^16..2^19 rows: their 2^15-row sub-blocks go through the same launch as virtual
    // columns, the stages >= 15 run in k_ntt_top (before the inverse launch / after the forward launch) through j.tmp.
    std::vector<size_t> order(jobs.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return jobs[x].log_n > jobs[y].log_n; });
    auto launch_top = [&](const LdeJob& j, size_t qi, bool fwd, cudaStream_t st) -> int {
        const uint32_t T = j.log_n - COL_MAX_LOG;                 // stages above the shared-memory kernel
        const uint32_t n_pass = (T + 3) / 4;
        std::vector<uint32_t> qs(n_pass, T / n_pass);
        for (uint32_t i = 0; i < T % n_pass; i++) qs[i]++;
        const size_t n = (size_t)1 << j.log_n, N = n << log_blowup;
        // forward: ascending from stage 15; inverse: descending from the top
        uint32_t jj = fwd ? COL_MAX_LOG : j.log_n;
        for (uint32_t pi = 0; pi < n_pass; pi++) {
            const uint32_t Q = qs[pi];
            if (!fwd) jj -= Q;
            const bool last_fwd = fwd && pi + 1 == n_pass, first_inv = !fwd && pi == 0;
            ColJob b = base;
            b.log_n = j.log_n;
            b.ctab = d_ctab + job_ctab[qi];
            b.src = first_inv ? j.src : j.tmp;
            b.src_col_stride = n;
            b.src_coset_stride = fwd ? (uint64_t)j.w * n : 0;
            if (last_fwd) {
                b.dst = j.dst;
                b.dst_col_stride = N;
                b.dst_coset_stride = n;
            } else {
                b.dst = j.tmp;
                b.dst_col_stride = n;
                b.dst_coset_stride = fwd ? (uint64_t)j.w * n : 0;
            }
            dim3 grid((unsigned)((n >> Q) / 256), j.w, fwd ? n_cosets : 1);
#define P3R_TOP(QQ)                                                                       \
    if (fwd) k_ntt_top<F, QQ, true><<<grid, 256, 0, st>>>(b, jj, last_fwd ? 1u : 0u);    \
    else k_ntt_top<F, QQ, false><<<grid, 256, 0, st>>>(b, jj, 0u)
            if (Q == 1) { P3R_TOP(1); }
            else if (Q == 2) { P3R_TOP(2); }
            else if (Q == 3) { P3R_TOP(3); }
            else { P3R_TOP(4); }
#undef P3R_TOP
            LAUNCH_CHECK_C(KC_NTT);
            if (fwd) jj += Q;
        }
        return P3R_OK;
    };
    // The kernel holds 128 KB of shared memory, so one CTA per SM: a launch of C CTAs takes ceil(C / 148) waves of ~30 us, and the
    // inverse launch of a commit (173 CTAs for the layer's main traces: two waves, the second 17 % full) is followed by a forward
    // launch that cannot start before it ends. The jobs are therefore split into up to N_AUX groups of similar size, each with its
    // own inverse -> forward chain on its own stream: a group's forward CTAs fill the SMs another group's inverse tail leaves idle.
    // Grouping: columns of up to 2^14 rows form their own group and run in CTAs of 2^14 elements (256 threads, 64 KB): two of
    // those fit an SM, so one CTA's copy-in / copy-out overlaps the other's butterflies, which a single 2^15-element CTA per SM
    // cannot do. Taller columns need the 2^15-element CTA. Within a size class the jobs are spread over the remaining streams.
    std::vector<std::vector<size_t>> groups;
    std::vector<uint32_t> group_cta_log;
    {
        std::vector<size_t> small, large;
        for (size_t qi : order) (ctx->lde_small_cta && jobs[qi].log_n <= COL_MAX_LOG - 1 ? small : large).push_back(qi);
        auto spread = [&](const std::vector<size_t>& js, uint32_t n, uint32_t cta_log) {
            if (js.empty()) return;
            n = (uint32_t)std::max<size_t>(1, std::min<size_t>(n, js.size()));
            const size_t g0 = groups.size();
            std::vector<uint64_t> load(n, 0);
            groups.resize(g0 + n);
            group_cta_log.resize(g0 + n, cta_log);
            for (size_t qi : js) {   // largest first, to the lightest group
                size_t g = std::min_element(load.begin(), load.end()) - load.begin();
                groups[g0 + g].push_back(qi);
                load[g] += (uint64_t)jobs[qi].w << jobs[qi].log_n;
            }
        };
        const uint32_t streams = ctx->lde_streams;
        const uint32_t n_small = small.empty() ? 0 : std::max(1u, streams / 2), n_large = std::max(1u, streams - n_small);
        spread(large, n_large, COL_MAX_LOG);
        spread(small, std::max(1u, n_small), COL_MAX_LOG - 1);
    }
    const bool multi = groups.size() > 1;
    // descriptors first (their copies are queued on the main stream), then the fork, then the launches on the group streams
    std::vector<const ColJob*> d_levels(groups.size() * 2, nullptr);
    std::vector<uint32_t> n_levels(groups.size() * 2, 0), n_ctas(groups.size() * 2, 0);
    for (size_t g = 0; g < groups.size(); g++) {
        for (int dir = 0; dir < 2; dir++) {
            std::vector<ColJob> level;
            uint32_t cta = 0;
            for (size_t qi : groups[g]) {
                const LdeJob& j = jobs[qi];
                const size_t n = (size_t)1 << j.log_n, N = n << log_blowup;
                const bool big = j.log_n > COL_MAX_LOG;
                const uint32_t sub_log = big ? COL_MAX_LOG : j.log_n;   // rows handled inside one CTA column
                const size_t sub_n = (size_t)1 << sub_log;
                ColJob b = base;
                b.log_n = sub_log;
                b.n_cols = j.w << (j.log_n - sub_log);                   // virtual columns = sub-blocks of 2^15 rows
                b.cols_per_cta = 1u << (group_cta_log[g] - sub_log);
                col_plan(sub_log, b.q);
                b.n_inv = finv<F>(to_monty<F>(1u << j.log_n));
                if (dir == 0) {
                    b.src = big ? j.tmp : j.src;
                    b.dst = j.coef;
                    b.src_col_stride = b.dst_col_stride = sub_n;         // == n for ordinary jobs; sub-blocks are contiguous
                    b.dst_coset_stride = 0;
                    b.n_cosets = 1;
                    b.ctab = d_ctab;
                } else {
                    b.src = j.coef;
                    b.src_col_stride = sub_n;
                    b.n_cosets = n_cosets;
                    b.ctab = d_ctab + job_ctab[qi];
                    if (big) {
                        b.dst = j.tmp;                                   // [coset][column][n], natural order
                        b.dst_col_stride = sub_n;
                        b.dst_coset_stride = (uint64_t)j.w * n;
                        b.natural_out = 1;
                    } else {
                        b.dst = j.dst;
                        b.dst_col_stride = N;
                        b.dst_coset_stride = n;
                    }
                }
                b.cta_begin = cta;
                cta += ((b.n_cols + b.cols_per_cta - 1) / b.cols_per_cta) * b.n_cosets;
                level.push_back(b);
            }
            const ColJob* d_jobs = upload_vec(ctx, level);
            if (!d_jobs) {
                set_err(ctx, "staging exhausted");
                return P3R_ERR_OOM;
            }
            d_levels[2 * g + dir] = d_jobs;
            n_levels[2 * g + dir] = (uint32_t)level.size();
            n_ctas[2 * g + dir] = cta;
        }
    }
    if (multi) TRY(fork_streams(ctx));
    for (size_t g = 0; g < groups.size(); g++) {
        cudaStream_t st = multi ? ctx->aux[g % p3r_ctx::N_AUX] : ctx->stream;
        const size_t smem = (size_t)4 << group_cta_log[g];
        const uint32_t threads = group_cta_log[g] == COL_MAX_LOG ? 512u : 256u;
        for (size_t qi : groups[g])
            if (jobs[qi].log_n > COL_MAX_LOG) TRY(launch_top(jobs[qi], qi, false, st));
        k_ntt_col<F, false><<<n_ctas[2 * g], threads, smem, st>>>(d_levels[2 * g], n_levels[2 * g]);
        LAUNCH_CHECK_C(KC_NTT);
        k_ntt_col<F, true><<<n_ctas[2 * g + 1], threads, smem, st>>>(d_levels[2 * g + 1], n_levels[2 * g + 1]);
        LAUNCH_CHECK_C(KC_NTT);
        for (size_t qi : groups[g])
            if (jobs[qi].log_n > COL_MAX_LOG) TRY(launch_top(jobs[qi], qi, true, st));
    }
    if (multi) TRY(join_streams(ctx));
    return P3R_OK;
}

// Batched coset LDE of several matrices (all tables of one commit round): the same pass level of every job shares a launch.
template <class F>
static int coset_lde_batch(p3r_ctx* ctx, const std::vector<LdeJob>& all_jobs, uint32_t log_blowup) {
    if (all_jobs.empty()) return P3R_OK;
    uint32_t max_logN = 0;
    uint64_t bytes = 0;
    std::vector<LdeJob> jobs, col_jobs;  // multi-pass tile kernel / whole-column kernel
    for (auto& j : all_jobs) {
        if (!j.w) continue;
        max_logN = std::max(max_logN, j.log_n + log_blowup);
        bytes += 4ull * (((size_t)1 << j.log_n) + ((size_t)1 << (j.log_n + log_blowup))) * j.w;
        (ctx->use_col_ntt && j.log_n >= COL_MIN_LOG && j.log_n <= COL_TOP_MAX_LOG ? col_jobs : jobs).push_back(j);
    }
    TRY(ensure_twiddles<F>(ctx, std::max(max_logN, 16u)));
    KT kt(ctx, KC_NTT, bytes);  // algorithmic bytes: read every trace once, write every LDE once
    TRY(coset_lde_cols<F>(ctx, col_jobs, log_blowup));
    if (jobs.empty()) return P3R_OK;
    std::vector<std::vector<PassPlan>> plans;
    std::vector<NttPass> base;
    size_t max_passes = 0;
    for (auto& j : jobs) {
        GTable gt;
        TRY(get_gtable<F>(ctx, j.log_n, &gt));
        NttPass a{};
        a.log_n = j.log_n;
        a.tw = ctx->tw;
        a.tws = ctx->tws;
        a.logT = ctx->logT;
        a.r4 = ctx->r4;
        a.r8 = ctx->r8;
        a.r8_3 = ctx->r8_3;
        a.g_lo = gt.lo;
        a.g_hi = gt.hi;
        a.use_g = j.use_g;
        a.rot = j.rot;
        a.n_inv = finv<F>(to_monty<F>(1u << j.log_n));
        a.log_blowup = log_blowup;
        a.n_cols = j.w;
        base.push_back(a);
        plans.push_back(plan_passes(j.log_n));
        max_passes = std::max(max_passes, plans.back().size());
    }
    // inverse: DIF, stages descending => every job walks its plan from the last pass to the first
    for (size_t lvl = 0; lvl < max_passes; lvl++) {
        std::vector<NttPass> level;
        for (size_t q = 0; q < jobs.size(); q++) {
            auto& plan = plans[q];
            if (lvl >= plan.size()) continue;
            size_t pi = plan.size() - 1 - lvl;
            size_t n = (size_t)1 << jobs[q].log_n;
            NttPass b = base[q];
            b.forward = 0;
            b.first = (lvl == 0);
            b.last = (pi == 0);
            b.s0 = plan[pi].s0;
            b.r = plan[pi].r;
            b.log_cw = plan[pi].log_cw;
            b.src = b.first ? jobs[q].src : jobs[q].coef;
            b.dst = jobs[q].coef;
            b.src_col_stride = b.dst_col_stride = n;
            b.dst_coset_stride = 0;
            b.n_cosets = 1;
            level.push_back(b);
        }
        TRY(launch_pass_level<F>(ctx, level));
    }
    // forward: DIT, stages ascending, all cosets of all jobs in one launch per level
    for (size_t lvl = 0; lvl < max_passes; lvl++) {
        std::vector<NttPass> level;
        for (size_t q = 0; q < jobs.size(); q++) {
            auto& plan = plans[q];
            if (lvl >= plan.size()) continue;
            size_t n = (size_t)1 << jobs[q].log_n, N = n << log_blowup;
            NttPass b = base[q];
            b.forward = 1;
            b.first = (lvl == 0);
            b.last = (lvl == plan.size() - 1);
            b.s0 = plan[lvl].s0;
            b.r = plan[lvl].r;
            b.log_cw = plan[lvl].log_cw;
            b.src = b.first ? jobs[q].coef : jobs[q].tmp;
            b.dst = b.last ? jobs[q].dst : jobs[q].tmp;
            b.src_col_stride = b.first ? n : N;
            b.dst_col_stride = N;
            b.dst_coset_stride = n;
            b.n_cosets = 1u << log_blowup;
            level.push_back(b);
        }
        TRY(launch_pass_level<F>(ctx, level));
    }
    return P3R_OK;
}
template <class F>
static int coset_lde(p3r_ctx* ctx, const uint32_t* src, uint32_t* dst, uint32_t log_n, uint32_t w, uint32_t log_blowup,
                     bool use_g, uint32_t rot, uint32_t* scratch_coef, uint32_t* scratch_tmp) {
    if (w == 0) return P3R_OK;
    return coset_lde_batch<F>(ctx, {LdeJob{src, dst, log_n, w, use_g, rot, scratch_coef, scratch_tmp}}, log_blowup);
}

// ------------------------------------------------------------------------------------------------
// K5 host: MerkleTreeMmcs::commit over column-major device matrices (mixed heights).
// ------------------------------------------------------------------------------------------------
// Every compression level of one tree (and, for row-major leaves, the leaf level). Levels with more than STAGE_MAX_NODES
// nodes are throughput-bound and use the one-thread-per-permutation kernel (one launch per level); the rest is latency-bound
// and runs as fused k_merkle_stage launches of up to STAGE_MAX_LEVELS levels each. `inj_at(level)` = device digests of the
// rows injected at that level (nullptr = none). `leaf_rows` != nullptr: level 0 is still to be hashed from that row-major
// matrix of `leaf_w` words per row; otherwise level 0 is already in `digests`.
template <class F, class InjAt>
static int build_tree(p3r_ctx* ctx, const uint32_t* leaf_rows, uint32_t leaf_w, uint32_t lmax, uint32_t* digests, InjAt inj_at) {
    const uint32_t cap = ctx->fri.cap_height;
    const uint32_t last_level = lmax - cap;
    const uint32_t rows = 1u << lmax;
    auto level_off = [&](uint32_t l) { return (((size_t)2 << lmax) - ((size_t)2 << (lmax - l))) * 8; };
    ctx->kstats.bytes[KC_COMPRESS] += 96ull * rows;
    ctx->kstats.perms[KC_COMPRESS] += ((uint64_t)rows - ((uint64_t)1 << cap));   // one compression per internal node
    if (leaf_rows) ctx->kstats.perms[rows > STAGE_MAX_NODES ? KC_HASH : KC_COMPRESS] += (uint64_t)rows * ((leaf_w + 7) / 8);
    bool leaves_done = leaf_rows == nullptr;
    if (!leaves_done && ctx->d_p2w) {   // width-24 leaf hasher: always its own launch (the fused stage kernel is width 16)
        KT kt(ctx, KC_HASH, (uint64_t)rows * (4ull * leaf_w + 32));
        k_hash_rows_rowmajor_w24<F><<<(rows + 127) / 128, 128, 0, ctx->stream>>>(leaf_rows, leaf_w, rows, digests, ctx->d_p2w);
        LAUNCH_CHECK_C(KC_HASH);
        leaves_done = true;
    }
    if (!leaves_done && rows > STAGE_MAX_NODES) {
        KT kt(ctx, KC_HASH, (uint64_t)rows * (4ull * leaf_w + 32));
        k_hash_rows_rowmajor<F><<<(rows + 127) / 128, 128, 0, ctx->stream>>>(leaf_rows, leaf_w, rows, digests);
        LAUNCH_CHECK_C(KC_HASH);
        leaves_done = true;
    }
    KT kt_tree(ctx, KC_COMPRESS, 0);
    uint32_t cur = 0;  // levels 0..cur are done (level 0 only if leaves_done)
    while (cur < last_level || !leaves_done) {
        const uint32_t l = cur + 1;
        const uint32_t n_next = l <= lmax ? 1u << (lmax - l) : 0;
        if (leaves_done && n_next > STAGE_MAX_NODES) {
            const uint32_t* inj = inj_at(l);
            if (inj) ctx->kstats.bytes[KC_COMPRESS] += 32ull * n_next, ctx->kstats.perms[KC_COMPRESS] += n_next;
            // one thread per permutation. (A one-level kernel with 16 lanes per node for the 2^13..2^15-node levels was measured
            // and dropped: compress class 1.03 -> 1.13 / 1.23 ms with it up to 2^14 / 2^15 nodes — the cooperative permutation has
            // an eighth of the throughput, and these levels already need it.)
            k_compress<F><<<(n_next + 127) / 128, 128, 0, ctx->stream>>>(digests + level_off(l - 1), digests + level_off(l), n_next, inj);
            LAUNCH_CHECK_C(KC_COMPRESS);
            cur = l;
            continue;
        }
        MerkleStage st{};
        st.digests = digests;
        st.log_max_h = lmax;
        st.first_level = l;
        // levels per launch: tunable (env P3R_STAGE_LEVELS, default below). Fewer levels per CTA = more CTAs, less issue
        // contention at the widest level, at the price of one more launch per tree.
        static const uint32_t stage_levels = [] {
            const char* e = getenv("P3R_STAGE_LEVELS");
            uint32_t v = e ? (uint32_t)atoi(e) : STAGE_DEFAULT_LEVELS;
            return std::max(1u, std::min(v, STAGE_MAX_LEVELS));
        }();
        st.n_levels = std::min(stage_levels, last_level - cur);
        st.with_leaves = leaves_done ? 0 : 1;
        st.leaf_rows = leaf_rows;
        st.leaf_w = leaf_w;
        if (st.with_leaves) ctx->kstats.bytes[KC_HASH] += (uint64_t)rows * (4ull * leaf_w + 32);
        for (uint32_t j = 0; j < st.n_levels; j++) {
            st.inj[j] = inj_at(l + j);
            if (st.inj[j]) ctx->kstats.bytes[KC_COMPRESS] += 32ull << (lmax - l - j), ctx->kstats.perms[KC_COMPRESS] += 1ull << (lmax - l - j);
        }
        const uint32_t grid = 1u << (lmax - (cur + st.n_levels));            // nodes of the stage's last level
        const uint32_t widest = st.with_leaves ? (1u << st.n_levels) : (1u << (st.n_levels - 1));
        const uint32_t threads = std::max(32u, std::min(widest * 16, 1024u));
        k_merkle_stage<F><<<grid, threads, 0, ctx->stream>>>(st, ctx->d_p2);
        LAUNCH_CHECK_C(KC_COMPRESS);
        leaves_done = true;
        cur += st.n_levels;
    }
    return P3R_OK;
}

// ------------------------------------------------------------------------------------------------
// K5 host: MerkleTreeMmcs::commit over column-major device matrices (mixed heights).
// ------------------------------------------------------------------------------------------------
template <class F>
static int commit_tree(p3r_ctx* ctx, const std::vector<MatRef>& mats, Tree* t, uint32_t* digests) {
    t->mats = mats;
    uint32_t lmax = 0;
    for (auto& m : mats) lmax = std::max(lmax, m.log_h);
    t->log_max_h = lmax;
    t->digests = digests;
    if (lmax < ctx->fri.cap_height) {
        set_err(ctx, "matrix shorter than the Merkle cap");
        return P3R_ERR_INVALID_ARG;
    }
    // one k_hash_rows launch for the rows of every height: level 0 for the tallest, injected digests for the others
    std::vector<HashJob> jobs;
    std::vector<const uint32_t*> inj_digests(lmax + 1, nullptr);  // by level
    uint64_t hash_bytes = 0;
    for (uint32_t lh = lmax + 1; lh-- > 0;) {
        std::vector<const uint32_t*> cols;
        for (auto& m : mats)
            if (m.log_h == lh)
                for (uint32_t c = 0; c < m.w; c++) cols.push_back(m.d + ((size_t)c << m.log_h));
        if (cols.empty()) continue;
        if (lh < ctx->fri.cap_height) {
            set_err(ctx, "matrix shorter than the Merkle cap");
            return P3R_ERR_INVALID_ARG;
        }
        HashJob j{};
        j.colptr = upload_vec(ctx, cols);
        j.ncols = (uint32_t)cols.size();
        j.n_rows = 1u << lh;
        j.out = lh == lmax ? digests : arena_alloc<uint32_t>(ctx, (size_t)8 << lh);
        if (!j.colptr || !j.out) {
            set_err(ctx, "staging / arena exhausted");
            return P3R_ERR_OOM;
        }
        if (lh != lmax) inj_digests[lmax - lh] = j.out;
        hash_bytes += (uint64_t)j.n_rows * (4ull * j.ncols + 32);
        ctx->kstats.perms[KC_HASH] += (uint64_t)j.n_rows * ((j.ncols + 7) / 8);
        jobs.push_back(j);
    }
    std::stable_sort(jobs.begin(), jobs.end(), [](const HashJob& a, const HashJob& b) { return a.ncols > b.ncols; });
    if (ctx->use_hash_queue && !ctx->d_p2w) {
        // work queue: items of 32 rows, longest sponges first, taken by the warps of a machine-filling grid
        uint32_t items = 0;
        for (auto& j : jobs) {
            j.cta_begin = items;
            items += (j.n_rows + 31) / 32;
        }
        std::vector<uint8_t> blob(sizeof(HashQueue) + jobs.size() * sizeof(HashJob));
        HashQueue hq{(uint32_t)jobs.size(), items, 0u, 0u};
        std::memcpy(blob.data(), &hq, sizeof hq);
        std::memcpy(blob.data() + sizeof hq, jobs.data(), jobs.size() * sizeof(HashJob));
        char* d_blob = (char*)upload_small(ctx, blob.data(), blob.size(), /*device_mutable=*/true);
        if (!d_blob) {
            set_err(ctx, "staging exhausted");
            return P3R_ERR_OOM;
        }
        const uint32_t grid = std::min((items + 3) / 4, ctx->n_sms * 16u);
        KT kt(ctx, KC_HASH, hash_bytes);
        k_hash_rows_queue<F><<<grid, 128, 0, ctx->stream>>>(reinterpret_cast<const HashJob*>(d_blob + sizeof hq),
                                                          reinterpret_cast<HashQueue*>(d_blob));
        LAUNCH_CHECK_C(KC_HASH);
    } else {
        // CTA size (P3R_HASH_CTA = 32 / 64 / 128): the long-sponge CTAs all land in the first wave, round-robin over the SMs, so
        // small CTAs spread them evenly. Measured per layer proof (hash class): 128 rows 1.18 ms, 64 rows 1.18 ms, 32 rows 1.13 ms.
        static const uint32_t hash_cta = [] {
            const char* e = getenv("P3R_HASH_CTA");
            uint32_t v = e ? (uint32_t)atoi(e) : 32u;
            return (v == 32 || v == 64 || v == 128) ? v : 32u;
        }();
        uint32_t cta = 0;
        for (auto& j : jobs) {
            j.cta_begin = cta;
            cta += (j.n_rows + hash_cta - 1) / hash_cta;
        }
        const HashJob* d_jobs = upload_vec(ctx, jobs);
        if (!d_jobs) {
            set_err(ctx, "staging exhausted");
            return P3R_ERR_OOM;
        }
        KT kt(ctx, KC_HASH, hash_bytes);
        if (ctx->d_p2w) k_hash_rows_w24<F><<<cta, hash_cta, 0, ctx->stream>>>(d_jobs, (uint32_t)jobs.size(), ctx->d_p2w);
        else k_hash_rows<F><<<cta, hash_cta, 0, ctx->stream>>>(d_jobs, (uint32_t)jobs.size());
        LAUNCH_CHECK_C(KC_HASH);
    }
    return build_tree<F>(ctx, nullptr, 0, lmax, digests, [&](uint32_t level) { return inj_digests[level]; });
}
static size_t tree_digest_words(uint32_t log_max_h) { return ((size_t)2 << log_max_h) * 8; }
static int read_cap(p3r_ctx* ctx, const Tree& t, uint32_t* cap_out) {
    uint32_t cap = ctx->fri.cap_height;
    uint32_t l = t.log_max_h - cap;
    CUDA_TRY(d2h_async(ctx, cap_out, t.digests + t.level_off(l) * 8, ((size_t)8 << cap) * 4));
    CUDA_TRY(ctx_wait(ctx));
    return P3R_OK;
}

// Upload a row-major host matrix and transpose to column-major device memory.
static int upload_matrix(p3r_ctx* ctx, const p3r_matrix_u32& m, uint32_t* d_rowmajor_scratch, uint32_t* d_colmajor) {
    size_t words = (size_t)m.height * m.width;
    if (!words) return P3R_OK;
    CUDA_TRY(cudaMemcpyAsync(d_rowmajor_scratch, m.data, words * 4, cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid((m.height + 31) / 32, (m.width + 31) / 32), block(32, 8);
    {
        KT kt(ctx, KC_TRANSPOSE, 8ull * words);
        k_transpose_in<<<grid, block, 0, ctx->stream>>>(d_rowmajor_scratch, d_colmajor, m.height, m.width);
        LAUNCH_CHECK_C(KC_TRANSPOSE);
    }
    return P3R_OK;
}

// K3 host: generate the Poseidon2 table of instance `d` on the device from its operation list.
template <class F>
static int fill_poseidon2_table(p3r_ctx* ctx, const InstDev& d, const p3r_poseidon2_ops& ops, uint32_t* d_colmajor) {
    const uint32_t H = 1u << d.log_h;
    constexpr uint32_t R = (F::SBOX == 7) ? 1 : 0;
    const uint32_t width = 16 + 8 * (16 * R + 16) + F::ROUNDS_P * (R + 1) + 2;
    if (d.main_w != width || d.prep_w != 24 || !d.prep_trace || ops.n_ops > H || !ops.input_values || !ops.mmcs_bit ||
        !ops.mmcs_index_sum) {
        set_err(ctx, "poseidon2 ops given for an instance that is not a Poseidon2 table of this field");
        return P3R_ERR_INVALID_ARG;
    }
    size_t n = ops.n_ops;
    uint32_t* d_in = arena_alloc<uint32_t>(ctx, std::max<size_t>(n * 16, 4));
    uint32_t* d_sum = arena_alloc<uint32_t>(ctx, std::max<size_t>(n, 4));
    uint8_t* d_bit = arena_alloc<uint8_t>(ctx, std::max<size_t>(n, 16));
    if (!d_in || !d_sum || !d_bit) return P3R_ERR_OOM;
    CUDA_TRY(cudaMemcpyAsync(d_in, ops.input_values, n * 64, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_sum, ops.mmcs_index_sum, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_bit, ops.mmcs_bit, n, cudaMemcpyHostToDevice, ctx->stream));
    P2FillArgs a{};
    a.inputs = d_in;
    a.mmcs_bit = d_bit;
    a.idx_sum = d_sum;
    a.new_start = d.prep_trace + (size_t)22 * H;    // preprocessed tail: [.., mmcs_idx, mmcs_flag, new_start, merkle_path]
    a.merkle_path = d.prep_trace + (size_t)23 * H;
    a.n_ops = ops.n_ops;
    a.log_h = d.log_h;
    a.out = d_colmajor;
    KT kt(ctx, KC_MISC, (uint64_t)H * width * 4);
    k_poseidon2_table_fill<F><<<(H + 127) / 128, 128, 0, ctx->stream>>>(a);
    LAUNCH_CHECK_C(KC_MISC);
    return P3R_OK;
}

// Host: generate the ALU table of instance `d` on the device from its schedule + operand values.
template <class F>
static int fill_alu_table(p3r_ctx* ctx, const InstDev& d, const p3r_alu_ops& ops, uint32_t* d_colmajor) {
    const uint32_t H = 1u << d.log_h;
    const uint32_t num_int = ops.k_max ? (ops.k_max - 1) / 2 : 0;
    if (ops.d != 4 || ops.k_max != 4 || ops.lanes == 0 ||
        d.main_w != ops.lanes * 16 + (num_int + 2 * (ops.k_max - 1) + 1) * 4 || (uint64_t)ops.n_slots > (uint64_t)H * ops.lanes ||
        (ops.n_slots && (!ops.slot_kind || !ops.slot_first)) || (ops.n_ops && !ops.values)) {
        set_err(ctx, "alu ops given for an instance that is not a D=4, k=4 ALU table of this shape");
        return P3R_ERR_INVALID_ARG;
    }
    for (uint32_t i = 0; i < ops.n_slots; i++) {   // every referenced operation must exist (the kernel does not re-check)
        const uint32_t k = ops.slot_kind[i];
        if (k > ops.k_max || (k && (uint64_t)ops.slot_first[i] + k > ops.n_ops) || (k >= 2 && i % ops.lanes != 0)) {
            set_err(ctx, "alu ops: malformed schedule slot " + std::to_string(i));
            return P3R_ERR_INVALID_ARG;
        }
    }
    const size_t ns = std::max<size_t>(ops.n_slots, 4), nv = std::max<size_t>((size_t)ops.n_ops * 16, 4);
    uint32_t* d_kind = arena_alloc<uint32_t>(ctx, ns);
    uint32_t* d_first = arena_alloc<uint32_t>(ctx, ns);
    uint32_t* d_val = arena_alloc<uint32_t>(ctx, nv);
    if (!d_kind || !d_first || !d_val) return P3R_ERR_OOM;
    CUDA_TRY(cudaMemcpyAsync(d_kind, ops.slot_kind, (size_t)ops.n_slots * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_first, ops.slot_first, (size_t)ops.n_slots * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_val, ops.values, (size_t)ops.n_ops * 64, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(d_colmajor, 0, (size_t)H * d.main_w * 4, ctx->stream));
    AluFillArgs a{};
    a.kind = d_kind;
    a.first = d_first;
    a.values = d_val;
    a.n_slots = ops.n_slots;
    a.lanes = ops.lanes;
    a.k_max = ops.k_max;
    a.log_h = d.log_h;
    a.wnr = ctx->w_m;
    a.out = d_colmajor;
    KT kt(ctx, KC_MISC, (uint64_t)H * d.main_w * 4);
    k_alu_table_fill<F><<<(H + 127) / 128, 128, 0, ctx->stream>>>(a);
    LAUNCH_CHECK_C(KC_MISC);
    return P3R_OK;
}
template <class F>
static int fill_from_ops(p3r_ctx* ctx, const InstDev& d, const p3r_table_ops& t, uint32_t* d_colmajor) {
    return t.poseidon2 ? fill_poseidon2_table<F>(ctx, d, *t.poseidon2, d_colmajor) : fill_alu_table<F>(ctx, d, *t.alu, d_colmajor);
}

// Cut a constraint program into QG_GROUPS sub-programs for k_quotient_grouped. Group g takes the constraints whose ASSERT lies in
// the g-th part of the instruction stream (the lowering emits a constraint's instructions just before its ASSERT, so equal
// instruction ranges are roughly equal work) and, by one backward liveness pass over the slot-allocated code, the instructions
// those constraints depend on: an instruction is kept iff the slot it writes is live (needed by a kept later instruction and
// not overwritten in between). Shared subexpressions are recomputed by every group that needs them.
static void slice_program(const p3r_insn* insns, uint32_t n, std::vector<p3r_insn>* out, uint32_t off[QG_GROUPS + 1]) {
    auto is_ext_dst = [](uint32_t op) { return op >= P3R_OP_E_PERM && op <= P3R_OP_E_SUBB; };
    out->clear();
    for (int g = 0; g < QG_GROUPS; g++) {
        const uint32_t lo = (uint32_t)((uint64_t)n * g / QG_GROUPS), hi = (uint32_t)((uint64_t)n * (g + 1) / QG_GROUPS);
        std::vector<uint8_t> live_b(MAX_B_SLOTS, 0), live_e(MAX_E_SLOTS, 0), keep(n, 0);
        for (uint32_t k = n; k-- > 0;) {
            const p3r_insn& in = insns[k];
            bool need = false;
            if (in.op == P3R_OP_ASSERT_B || in.op == P3R_OP_ASSERT_E) {
                need = k >= lo && k < hi;
            } else if (in.op == P3R_OP_OUT_B) {
                need = false;
            } else if (is_ext_dst(in.op)) {
                need = in.dst < (uint32_t)MAX_E_SLOTS && live_e[in.dst];
                if (need) live_e[in.dst] = 0;
            } else {
                need = in.dst < (uint32_t)MAX_B_SLOTS && live_b[in.dst];
                if (need) live_b[in.dst] = 0;
            }
            if (!need) continue;
            keep[k] = 1;
            switch (in.op) {
                case P3R_OP_B_ADD: case P3R_OP_B_SUB: case P3R_OP_B_MUL: live_b[in.a] = live_b[in.b] = 1; break;
                case P3R_OP_B_NEG: case P3R_OP_ASSERT_B: case P3R_OP_E_FROMB: live_b[in.a] = 1; break;
                case P3R_OP_E_ADD: case P3R_OP_E_SUB: case P3R_OP_E_MUL: live_e[in.a] = live_e[in.b] = 1; break;
                case P3R_OP_E_NEG: case P3R_OP_ASSERT_E: live_e[in.a] = 1; break;
                case P3R_OP_E_MULB: case P3R_OP_E_ADDB: case P3R_OP_E_SUBB: live_e[in.a] = 1; live_b[in.b] = 1; break;
                default: break;   // leaves: no slot operands
            }
        }
        off[g] = (uint32_t)out->size();
        for (uint32_t k = 0; k < n; k++)
            if (keep[k]) out->push_back(insns[k]);
    }
    off[QG_GROUPS] = (uint32_t)out->size();
}

static bool is_pow2(uint32_t x) { return x && !(x & (x - 1)); }
static uint32_t ilog2(uint32_t x) {
    uint32_t l = 0;
    while ((1u << l) < x) l++;
    return l;
}

// ------------------------------------------------------------------------------------------------
// prep
// ------------------------------------------------------------------------------------------------
template <class F>
static int prep_commit_impl(p3r_ctx* ctx, uint32_t n_inst, const p3r_instance_desc* descs, const p3r_matrix_u32* prep,
                            p3r_prep** out, uint32_t* cap_out, uint32_t* has_prep_out) {
    auto* pp = new p3r_prep();
    pp->ctx = ctx;
    auto fail = [&](int rc) {
        for (void* p : pp->owned) cudaFree(p);
        delete pp;
        return rc;
    };
    // Device memory owned by the prep: sub-allocated from a few large slabs (cudaMalloc goes through the kernel-mode driver
    // and costs milliseconds when anything else holds its lock; ~80 separate allocations dominated this call).
    char* slab = nullptr;
    size_t slab_left = 0;
    auto dmalloc = [&](size_t bytes) -> void* {
        bytes = (std::max<size_t>(bytes, 16) + 255) & ~(size_t)255;
        if (bytes > slab_left) {
            size_t sz = std::max<size_t>(bytes, (size_t)32 << 20);
            void* p = nullptr;
            if (cudaMalloc(&p, sz) != cudaSuccess) return nullptr;
            pp->owned.push_back(p);
            slab = static_cast<char*>(p);
            slab_left = sz;
        }
        void* r = slab;
        slab += bytes;
        slab_left -= bytes;
        return r;
    };
    ctx->arena.reset();
    ring_reset(ctx);
    const bool trace_prep = getenv("P3R_TRACE_PREP") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!trace_prep) return;
        ctx_wait(ctx);
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[p3r prep] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    const uint32_t lb = ctx->fri.log_blowup;
    uint32_t max_logN = 0;
    for (uint32_t i = 0; i < n_inst; i++) max_logN = std::max(max_logN, descs[i].log_height + lb);
    {
        int rc = ensure_twiddles<F>(ctx, max_logN);
        if (rc) return fail(rc);
    }
    lap("twiddles");
    for (uint32_t i = 0; i < n_inst; i++) {
        const p3r_instance_desc& d = descs[i];
        InstDev s;
        s.log_h = d.log_height;
        s.main_w = d.main_width;
        s.prep_w = d.prep_width;
        s.n_pub = d.n_public;
        s.log_qc = d.log_quotient_chunks;
        s.uses_next = d.uses_next_row;
        if (d.log_quotient_chunks > lb) {
            set_err(ctx, "log_quotient_chunks exceeds log_blowup");
            return fail(P3R_ERR_INVALID_ARG);
        }
        if (d.constraints.n_base_slots > (uint32_t)MAX_B_SLOTS || d.constraints.n_ext_slots > (uint32_t)MAX_E_SLOTS ||
            d.lookup_inputs.n_base_slots > (uint32_t)MAX_B_SLOTS || d.lookup_inputs.n_outputs > (uint32_t)MAX_LK_OUTS ||
            d.n_interactions > (uint32_t)MAX_INTERACTIONS) {
            set_err(ctx, "program exceeds interpreter limits (slots/outputs/interactions)");
            return fail(P3R_ERR_UNSUPPORTED);
        }
        auto up = [&](const void* src, size_t bytes) -> void* {
            void* p = dmalloc(bytes);
            if (p && bytes) cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
            return p;
        };
        {
            // FNV-1a over the instruction words: selects a build-time specialised quotient kernel (specialized_gen.cuh)
            uint64_t h = 0xCBF29CE484222325ull;
            const uint8_t* bytes = reinterpret_cast<const uint8_t*>(d.constraints.insns);
            for (size_t k = 0; k < (size_t)d.constraints.n_insns * 16; k++) {
                h ^= bytes[k];
                h *= 0x100000001B3ull;
            }
            size_t n_spec = 0;
            const SpecEntry* reg = p3r_spec_registry(&n_spec);
            for (size_t q = 0; q < n_spec; q++)
                if (reg[q].hash == h && reg[q].field_id == ctx->field_id && reg[q].n_insns == d.constraints.n_insns) s.spec = reg[q].fn;
        }
        s.n_cons_insns = d.constraints.n_insns;
        s.n_constraints = d.constraints.n_constraints;
        s.cons = (uint4*)up(d.constraints.insns, (size_t)d.constraints.n_insns * 16);
        s.cons_econst = (Ext4*)up(d.constraints.ext_consts, (size_t)d.constraints.n_ext_consts * 16);
        if (d.constraints.n_insns >= 256) {
            std::vector<p3r_insn> sliced;
            slice_program(d.constraints.insns, d.constraints.n_insns, &sliced, s.goff);
            s.gcons = (uint4*)up(sliced.data(), sliced.size() * sizeof(p3r_insn));
            if (!s.gcons) {
                set_err(ctx, "device allocation failed");
                return fail(P3R_ERR_OOM);
            }
        }
        s.n_lk_insns = d.lookup_inputs.n_insns;
        s.n_lk_outs = d.lookup_inputs.n_outputs;
        s.lk = (uint4*)up(d.lookup_inputs.insns, (size_t)d.lookup_inputs.n_insns * 16);
        s.lookups.assign(d.lookups, d.lookups + d.n_lookups);
        s.inter.assign(d.interactions, d.interactions + d.n_interactions);
        s.d_lookups = (p3r_lookup*)up(d.lookups, (size_t)d.n_lookups * sizeof(p3r_lookup));
        s.d_inter = (p3r_interaction*)up(d.interactions, (size_t)d.n_interactions * sizeof(p3r_interaction));
        if (d.n_lookups && d.lookup_inputs.n_insns) {
            // FNV-1a over the lookup-input program and the lookup / interaction structure (scripts/gen_specialized.py logup_hash)
            uint64_t h = 0xCBF29CE484222325ull;
            auto mix = [&](uint32_t w) {
                for (int k = 0; k < 4; k++) {
                    h ^= (w >> (8 * k)) & 0xFFu;
                    h *= 0x100000001B3ull;
                }
            };
            const uint32_t* iw = reinterpret_cast<const uint32_t*>(d.lookup_inputs.insns);
            for (size_t k = 0; k < (size_t)d.lookup_inputs.n_insns * 4; k++) mix(iw[k]);
            for (uint32_t k = 0; k < d.n_lookups; k++) mix(d.lookups[k].first_interaction), mix(d.lookups[k].n_interactions);
            for (uint32_t k = 0; k < d.n_interactions; k++)
                mix(d.interactions[k].mult_out), mix(d.interactions[k].elem_out_first), mix(d.interactions[k].n_elems);
            size_t n_spec = 0;
            const SpecLogupEntry* reg = p3r_spec_logup_registry(&n_spec);
            for (size_t q = 0; q < n_spec; q++)
                if (reg[q].hash == h && reg[q].field_id == ctx->field_id && reg[q].n_insns == d.lookup_inputs.n_insns)
                    s.spec_logup = reg[q].fn;
        }
        if (!s.cons || !s.cons_econst || !s.lk || !s.d_lookups || !s.d_inter) {
            set_err(ctx, "device allocation failed");
            return fail(P3R_ERR_OOM);
        }
        lap("programs");
        for (auto& it : s.inter) pp->max_msg_w = std::max(pp->max_msg_w, it.n_elems);
        for (auto& l : s.lookups) pp->n_buses = std::max(pp->n_buses, l.bus + 1);
        if (pp->max_msg_w > 8) {
            set_err(ctx, "lookup tuples wider than 8 are not supported");
            return fail(P3R_ERR_UNSUPPORTED);
        }
        if (!s.lookups.empty()) pp->has_perm = true;
        // selectors on the quotient domain
        uint32_t NQ = 1u << (s.log_h + s.log_qc);
        s.sel = (uint32_t*)dmalloc((size_t)3 * NQ * 4);
        s.inv_van = (uint32_t*)dmalloc(((size_t)4 << s.log_qc));
        if (!s.sel || !s.inv_van) return fail(P3R_ERR_OOM);
        k_selectors<F><<<(NQ + 255) / 256, 256, 0, ctx->stream>>>(s.sel, s.inv_van, s.log_h, s.log_qc, ctx->gen_m, ctx->tw,
                                                                  ctx->logT);
        ctx->launches++;
        lap("selectors");
        if (s.prep_w) {
            if (!prep || !prep[i].data || prep[i].height != (1u << s.log_h) || prep[i].width != s.prep_w) {
                set_err(ctx, "preprocessed matrix shape mismatch");
                return fail(P3R_ERR_INVALID_ARG);
            }
            pp->has_prep = true;
            size_t n = (size_t)1 << s.log_h;
            s.prep_trace = (uint32_t*)dmalloc(n * s.prep_w * 4);
            s.prep_lde = (uint32_t*)dmalloc((n << lb) * s.prep_w * 4);
            uint32_t* rm = arena_alloc<uint32_t>(ctx, n * s.prep_w);
            uint32_t* coef = arena_alloc<uint32_t>(ctx, n * s.prep_w);
            uint32_t* tmp = s.log_h > TILE_LOG ? arena_alloc<uint32_t>(ctx, (n << lb) * s.prep_w) : nullptr;
            if (!s.prep_trace || !s.prep_lde || !rm || !coef || (s.log_h > TILE_LOG && !tmp)) return fail(P3R_ERR_OOM);
            lap("alloc");
            int rc = upload_matrix(ctx, prep[i], rm, s.prep_trace);
            if (rc) return fail(rc);
            lap("upload");
            rc = coset_lde<F>(ctx, s.prep_trace, s.prep_lde, s.log_h, s.prep_w, lb, true, 0, coef, tmp);
            if (rc) return fail(rc);
            lap("lde");
        }
        pp->inst.push_back(std::move(s));
    }
    if (pp->has_prep) {
        std::vector<MatRef> mats;
        uint32_t lmax = 0;
        for (auto& s : pp->inst)
            if (s.prep_w) {
                mats.push_back({s.prep_lde, s.log_h + lb, s.prep_w});
                lmax = std::max(lmax, s.log_h + lb);
            }
        uint32_t* dg = (uint32_t*)dmalloc(tree_digest_words(lmax) * 4);
        if (!dg) return fail(P3R_ERR_OOM);
        int rc = commit_tree<F>(ctx, mats, &pp->prep_tree, dg);
        if (rc) return fail(rc);
        pp->prep_cap.resize((size_t)8 << ctx->fri.cap_height);
        rc = read_cap(ctx, pp->prep_tree, pp->prep_cap.data());
        if (rc) return fail(rc);
        if (cap_out) std::memcpy(cap_out, pp->prep_cap.data(), pp->prep_cap.size() * 4);
        lap("tree");
    }
    if (ctx_wait(ctx) != cudaSuccess) {
        set_err(ctx, "prep: stream sync failed");
        return fail(P3R_ERR_CUDA);
    }
    if (has_prep_out) *has_prep_out = pp->has_prep;
    *out = pp;
    return P3R_OK;
}

// ------------------------------------------------------------------------------------------------
// session phases
// ------------------------------------------------------------------------------------------------
template <class F>
static int prove_begin_impl(p3r_ctx* ctx, const p3r_prep* prep, const p3r_matrix_u32* traces,
                            const uint32_t* const* public_values, p3r_session** out, const p3r_traces* resident = nullptr,
                            const p3r_table_ops* tops = nullptr) {
    ctx->arena.reset();
    ring_reset(ctx);
    auto* s = new p3r_session();
    s->ctx = ctx;
    s->prep = prep;
    const uint32_t lb = ctx->fri.log_blowup;
    size_t n_inst = prep->inst.size();
    s->trace.assign(n_inst, nullptr);
    s->main_lde.assign(n_inst, nullptr);
    s->perm.assign(n_inst, nullptr);
    s->perm_lde.assign(n_inst, nullptr);
    s->chunks.assign(n_inst, nullptr);
    s->chunk_lde.assign(n_inst, nullptr);
    s->d_pub.assign(n_inst, nullptr);
    size_t max_rm = 0;
    for (size_t i = 0; i < n_inst; i++) {
        const InstDev& d = prep->inst[i];
        if (d.n_pub && (!public_values || !public_values[i])) {
            set_err(ctx, "public values missing for instance " + std::to_string(i) + " (n_public = " + std::to_string(d.n_pub) + ")");
            delete s;
            return P3R_ERR_INVALID_ARG;
        }
        const bool from_ops = tops && (tops[i].poseidon2 || tops[i].alu);
        if (!resident && !from_ops && (traces[i].height != (1u << d.log_h) || traces[i].width != d.main_w || !traces[i].data)) {
            set_err(ctx, "trace shape mismatch for instance " + std::to_string(i));
            delete s;
            return P3R_ERR_INVALID_ARG;
        }
        size_t n = (size_t)1 << d.log_h;
        if (!from_ops) max_rm = std::max(max_rm, n * d.main_w);
    }
    uint32_t* rm = resident ? reinterpret_cast<uint32_t*>(ctx->dstage) : arena_alloc<uint32_t>(ctx, max_rm);
    if (!rm) {
        set_err(ctx, "device allocation failed");
        delete s;
        return P3R_ERR_OOM;
    }
    for (size_t i = 0; i < n_inst; i++) {
        const InstDev& d = prep->inst[i];
        size_t n = (size_t)1 << d.log_h;
        s->trace[i] = resident ? resident->d[i] : arena_alloc<uint32_t>(ctx, n * d.main_w);
        s->main_lde[i] = arena_alloc<uint32_t>(ctx, (n << lb) * d.main_w);
        if (!s->trace[i] || !s->main_lde[i]) {
            set_err(ctx, "device allocation failed");
            delete s;
            return P3R_ERR_OOM;
        }
        if (!resident) {
            int rc = (tops && (tops[i].poseidon2 || tops[i].alu)) ? fill_from_ops<F>(ctx, d, tops[i], s->trace[i])
                                                                  : upload_matrix(ctx, traces[i], rm, s->trace[i]);
            if (rc) {
                delete s;
                return rc;
            }
        }
        if (d.n_pub) {
            s->d_pub[i] = (uint32_t*)upload_small(ctx, public_values[i], (size_t)d.n_pub * 4);
            if (!s->d_pub[i]) {
                set_err(ctx, "staging exhausted (public values)");
                delete s;
                return P3R_ERR_OOM;
            }
        } else {
            s->d_pub[i] = reinterpret_cast<uint32_t*>(ctx->dstage);
        }
    }
    *out = s;
    return P3R_OK;
}

template <class F>
static int commit_main_impl(p3r_session* s, uint32_t* cap_out) {
    p3r_ctx* ctx = s->ctx;
    if (s->phase != PH_BEGIN) {
        set_err(ctx, "commit_main: wrong phase");
        return P3R_ERR_STATE;
    }
    const uint32_t lb = ctx->fri.log_blowup;
    std::vector<MatRef> mats;
    uint32_t lmax = 0;
    std::vector<LdeJob> jobs;
    for (size_t i = 0; i < s->prep->inst.size(); i++) {
        const InstDev& d = s->prep->inst[i];
        size_t n = (size_t)1 << d.log_h;
        uint32_t* coef = arena_alloc<uint32_t>(ctx, n * d.main_w);
        uint32_t* tmp = d.log_h > TILE_LOG ? arena_alloc<uint32_t>(ctx, (n << lb) * d.main_w) : nullptr;
        if (!coef || (d.log_h > TILE_LOG && !tmp)) return P3R_ERR_OOM;
        jobs.push_back({s->trace[i], s->main_lde[i], d.log_h, d.main_w, true, 0, coef, tmp});
        mats.push_back({s->main_lde[i], d.log_h + lb, d.main_w});
        lmax = std::max(lmax, d.log_h + lb);
    }
    TRY(coset_lde_batch<F>(ctx, jobs, lb));
    uint32_t* dg = arena_alloc<uint32_t>(ctx, tree_digest_words(lmax));
    if (!dg) return P3R_ERR_OOM;
    TRY(commit_tree<F>(ctx, mats, &s->main_tree, dg));
    TRY(read_cap(ctx, s->main_tree, cap_out));
    s->phase = PH_MAIN;
    return P3R_OK;
}

template <class F>
static int commit_perm_impl(p3r_session* s, const uint32_t alpha[4], const uint32_t beta[4], uint32_t* cap_out,
                            uint32_t* terminals_out) {
    p3r_ctx* ctx = s->ctx;
    if (s->phase != PH_MAIN) {
        set_err(ctx, "commit_perm: wrong phase");
        return P3R_ERR_STATE;
    }
    const p3r_prep* pp = s->prep;
    size_t n_inst = pp->inst.size();
    s->d_terminals = arena_alloc<Ext4>(ctx, n_inst);
    s->chal_off.assign(n_inst, 0);
    if (!pp->has_perm) {
        s->d_chal = reinterpret_cast<Ext4*>(ctx->dstage);
        s->phase = PH_PERM;
        return P3R_OK;
    }
    const uint32_t lb = ctx->fri.log_blowup, wnr = ctx->w_m;
    Ext4 a, b;
    std::memcpy(a.c, alpha, 16);
    std::memcpy(b.c, beta, 16);
    // challenge layout (recursion/src/verifier/batch_stark.rs:1086-1110): gamma = beta^W, prefix[bus] = alpha + (bus+1)*gamma
    Ext4 gamma = b;
    for (uint32_t i = 1; i < pp->max_msg_w; i++) gamma = emul<F>(gamma, b, wnr);
    std::vector<Ext4> prefix(pp->n_buses);
    Ext4 pr = a;
    for (uint32_t i = 0; i < pp->n_buses; i++) {
        pr = eadd<F>(pr, gamma);
        prefix[i] = pr;
    }
    std::vector<Ext4> chal;
    for (size_t i = 0; i < n_inst; i++) {
        s->chal_off[i] = (uint32_t)chal.size();
        for (auto& l : pp->inst[i].lookups) {
            chal.push_back(prefix[l.bus]);
            chal.push_back(b);
        }
    }
    // coefficient of tuple element k in a denominator (p3r_conventions): s * beta^(first_power + k), or with the powers in
    // descending order over the tuple (then every tuple must have the same length, the widest one)
    const p3r_conventions& cv = ctx->conv;
    if (cv.logup_descending)
        for (auto& d : pp->inst)
            for (auto& it : d.inter)
                if (it.n_elems != pp->max_msg_w) {
                    set_err(ctx, "logup_descending needs every lookup tuple to have the same length");
                    return P3R_ERR_UNSUPPORTED;
                }
    std::vector<Ext4> pw(9), bp(8);
    pw[0] = ext_one<F>();
    for (int k = 1; k < 9; k++) pw[k] = emul<F>(pw[k - 1], b, wnr);
    for (uint32_t k = 0; k < 8; k++) {
        const uint32_t e = cv.logup_first_power + (cv.logup_descending ? (k < pp->max_msg_w ? pp->max_msg_w - 1 - k : 0) : k);
        bp[k] = pw[std::min(e, 8u)];
        if (cv.logup_negate) bp[k] = eneg<F>(bp[k]);
    }
    s->d_chal = upload_vec(ctx, chal);
    Ext4* d_bp = upload_vec(ctx, bp);
    if (!s->d_chal || !d_bp || !s->d_terminals) return P3R_ERR_OOM;
    std::vector<MatRef> mats;
    std::vector<LdeJob> jobs;
    std::vector<LogupArgs> logup_tables;
    std::vector<SpecLogupKernel> logup_spec;   // generated kernel per table, or nullptr (interpreter)
    uint64_t logup_bytes = 0;
    uint32_t lmax = 0;
    for (size_t i = 0; i < n_inst; i++) {
        const InstDev& d = pp->inst[i];
        if (d.lookups.empty()) continue;
        size_t n = (size_t)1 << d.log_h;
        uint32_t pw = d.aux_w() * 4;
        s->perm[i] = arena_alloc<uint32_t>(ctx, n * pw);
        s->perm_lde[i] = arena_alloc<uint32_t>(ctx, (n << lb) * pw);
        Ext4* rowsum = arena_alloc<Ext4>(ctx, n);
        if (!s->perm[i] || !s->perm_lde[i] || !rowsum) return P3R_ERR_OOM;
        LogupArgs la{};
        la.insns = d.lk;
        la.n_insns = d.n_lk_insns;
        la.main = s->trace[i];
        la.prep = d.prep_trace;
        la.pub = s->d_pub[i];
        la.log_n = d.log_h;
        la.lookups = d.d_lookups;
        la.n_lookups = (uint32_t)d.lookups.size();
        la.inter = d.d_inter;
        la.chal = s->d_chal + s->chal_off[i];
        la.beta_pows = d_bp;
        la.perm = s->perm[i];
        la.rowsum = rowsum;
        la.wnr = wnr;
        la.chunk_sum = arena_alloc<Ext4>(ctx, (n + SCAN_CHUNK - 1) / SCAN_CHUNK);
        la.terminal = s->d_terminals + i;
        if (!la.chunk_sum) return P3R_ERR_OOM;
        logup_tables.push_back(la);
        logup_spec.push_back(ctx->use_spec ? d.spec_logup : nullptr);
        logup_bytes += (uint64_t)n * 4 * (d.main_w + d.prep_w + pw);
        uint32_t* coef = arena_alloc<uint32_t>(ctx, n * pw);
        uint32_t* tmp = d.log_h > TILE_LOG ? arena_alloc<uint32_t>(ctx, (n << lb) * pw) : nullptr;
        if (!coef || (d.log_h > TILE_LOG && !tmp)) return P3R_ERR_OOM;
        jobs.push_back({s->perm[i], s->perm_lde[i], d.log_h, pw, true, 0, coef, tmp});
        mats.push_back({s->perm_lde[i], d.log_h + lb, pw});
        lmax = std::max(lmax, d.log_h + lb);
    }
    {
        // rows: tables with a generated kernel get their own launch (32 rows x lookup groups per CTA), the others share one
        // interpreter launch; chunk sums and scan: one launch each for all tables
        std::vector<size_t> order(logup_tables.size());
        for (size_t i = 0; i < order.size(); i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return logup_tables[x].log_n > logup_tables[y].log_n; });
        std::vector<LogupArgs> sorted_tabs, generic;
        std::vector<SpecLogupKernel> sorted_spec;
        uint32_t cta = 0, scan_cta = 0;
        for (size_t i : order) {
            LogupArgs la = logup_tables[i];
            const uint32_t n = 1u << la.log_n;
            la.scan_cta_begin = scan_cta;
            scan_cta += (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
            if (!logup_spec[i]) {
                la.cta_begin = cta;
                cta += (n + 127) / 128;
                generic.push_back(la);
            }
            sorted_tabs.push_back(la);
            sorted_spec.push_back(logup_spec[i]);
        }
        const LogupArgs* d_tabs = upload_vec(ctx, sorted_tabs);
        if (!d_tabs) return P3R_ERR_OOM;
        const LogupArgs* d_gen = generic.empty() ? nullptr : upload_vec(ctx, generic);
        if (!generic.empty() && !d_gen) return P3R_ERR_OOM;
        KT kt(ctx, KC_LOGUP, logup_bytes);
        const uint32_t nt = (uint32_t)sorted_tabs.size();
        TRY(fork_streams(ctx));   // independent per-table kernels on side streams
        for (size_t i = 0; i < sorted_tabs.size(); i++)
            if (sorted_spec[i]) {
                p3r_spec_logup_launch(sorted_spec[i], sorted_tabs[i], ((1u << sorted_tabs[i].log_n) + 31) / 32,
                                      ctx->aux[i % p3r_ctx::N_AUX]);
                LAUNCH_CHECK_C(KC_LOGUP);
            }
        if (d_gen) {
            k_logup_rows<F><<<cta, 128, 0, ctx->aux[p3r_ctx::N_AUX - 1]>>>(d_gen, (uint32_t)generic.size());
            LAUNCH_CHECK_C(KC_LOGUP);
        }
        TRY(join_streams(ctx));
        k_logup_chunk_sums<F><<<scan_cta, 256, 0, ctx->stream>>>(d_tabs, nt);
        LAUNCH_CHECK_C(KC_LOGUP);
        k_logup_scan_apply<F><<<scan_cta, 256, 0, ctx->stream>>>(d_tabs, nt);
        LAUNCH_CHECK_C(KC_LOGUP);
    }
    TRY(coset_lde_batch<F>(ctx, jobs, lb));
    uint32_t* dg = arena_alloc<uint32_t>(ctx, tree_digest_words(lmax));
    if (!dg) return P3R_ERR_OOM;
    TRY(commit_tree<F>(ctx, mats, &s->perm_tree, dg));
    // terminals of instances with lookups, in order
    std::vector<Ext4> term(n_inst);
    CUDA_TRY(d2h_async(ctx, term.data(), s->d_terminals, n_inst * sizeof(Ext4)));
    TRY(read_cap(ctx, s->perm_tree, cap_out));
    size_t k = 0;
    for (size_t i = 0; i < n_inst; i++)
        if (!pp->inst[i].lookups.empty()) {
            std::memcpy(terminals_out + 4 * k, term[i].c, 16);
            k++;
        }
    s->phase = PH_PERM;
    return P3R_OK;
}

template <class F>
static int commit_quotient_impl(p3r_session* s, const uint32_t alpha[4], uint32_t* cap_out) {
    p3r_ctx* ctx = s->ctx;
    if (s->phase != PH_PERM) {
        set_err(ctx, "commit_quotient: wrong phase");
        return P3R_ERR_STATE;
    }
    const p3r_prep* pp = s->prep;
    const uint32_t lb = ctx->fri.log_blowup, wnr = ctx->w_m;
    Ext4 al;
    std::memcpy(al.c, alpha, 16);
    std::vector<MatRef> mats;
    std::vector<LdeJob> jobs;
    uint32_t lmax = 0;
    // alpha^{N-1-k} for every table, eight tables per launch
    std::vector<Ext4*> alpha_pows(pp->inst.size());
    for (size_t i0 = 0; i0 < pp->inst.size(); i0 += 8) {
        PowDescJobs pj{};
        uint32_t max_n = 1, cnt = 0;
        for (size_t i = i0; i < std::min(pp->inst.size(), i0 + 8); i++, cnt++) {
            const uint32_t nc = pp->inst[i].n_constraints;
            alpha_pows[i] = arena_alloc<Ext4>(ctx, std::max<uint32_t>(nc, 1));
            if (!alpha_pows[i]) return P3R_ERR_OOM;
            pj.out[cnt] = alpha_pows[i];
            pj.n[cnt] = nc;
            max_n = std::max(max_n, nc);
        }
        k_ext_powers_desc<F><<<dim3((max_n + 127) / 128, cnt), 128, 0, ctx->stream>>>(pj, al, wnr);
        LAUNCH_CHECK();
    }
    {
    KT kt_quot(ctx, KC_QUOTIENT, 0);
    TRY(fork_streams(ctx));
    for (size_t i = 0; i < pp->inst.size(); i++) {
        const InstDev& d = pp->inst[i];
        size_t n = (size_t)1 << d.log_h;
        uint32_t qc = 1u << d.log_qc;
        s->chunks[i] = arena_alloc<uint32_t>(ctx, n * 4 * qc);
        s->chunk_lde[i] = arena_alloc<uint32_t>(ctx, (n << lb) * 4 * qc);
        Ext4* ap = alpha_pows[i];
        if (!s->chunks[i] || !s->chunk_lde[i]) return P3R_ERR_OOM;
        QuotientArgs qa{};
        qa.insns = d.cons;
        qa.n_insns = d.n_cons_insns;
        qa.main = s->main_lde[i];
        qa.prep = d.prep_lde;
        qa.perm = s->perm_lde[i];
        qa.pub = s->d_pub[i];
        qa.log_n = d.log_h;
        qa.log_qc = d.log_qc;
        qa.log_blowup = lb;
        qa.sel = d.sel;
        qa.inv_van = d.inv_van;
        qa.chal = s->d_chal + s->chal_off[i];
        qa.pval = s->d_terminals + i;
        qa.econst = d.cons_econst;
        qa.alpha_pows = ap;
        qa.chunks = s->chunks[i];
        qa.wnr = wnr;
        uint32_t NQ = (uint32_t)(n << d.log_qc);
        {
            // the tables' quotient kernels are independent and latency-bound: one side stream each (round robin)
            cudaStream_t qs = ctx->aux[i % p3r_ctx::N_AUX];
            ctx->kstats.bytes[KC_QUOTIENT] += (uint64_t)NQ * (8ull * (d.main_w + d.prep_w + d.aux_w() * 4) + 16);
            if (d.spec && ctx->use_spec)
                p3r_spec_launch(d.spec, qa, (NQ + 31) / 32, p3r_spec_threads(), qs);  // 32 rows x constraint groups per CTA
            else if (d.gcons && ctx->use_grouped_interp) {
                QuotientGroups gr{};
                gr.insns = d.gcons;
                for (int g = 0; g <= QG_GROUPS; g++) gr.off[g] = d.goff[g];
                k_quotient_grouped<F><<<(NQ + 31) / 32, 32 * QG_GROUPS, 0, qs>>>(qa, gr);
            } else
                k_quotient<F><<<(NQ + 127) / 128, 128, 0, qs>>>(qa);
            LAUNCH_CHECK_C(KC_QUOTIENT);
        }
        // chunk c lives on the coset GENERATOR * w_NQ^c * H_n: LDE without the GENERATOR factor, rotated by -c*(N/NQ)
        uint32_t logN = d.log_h + lb;
        for (uint32_t c = 0; c < qc; c++) {
            uint32_t rot = (uint32_t)((((uint64_t)1 << logN) - ((uint64_t)c << (lb - d.log_qc))) & (((uint64_t)1 << logN) - 1));
            uint32_t* src = s->chunks[i] + (size_t)c * 4 * n;
            uint32_t* dst = s->chunk_lde[i] + (size_t)c * 4 * (n << lb);
            uint32_t* coef = arena_alloc<uint32_t>(ctx, n * 4);
            uint32_t* tmp = d.log_h > TILE_LOG ? arena_alloc<uint32_t>(ctx, (n << lb) * 4) : nullptr;
            if (!coef || (d.log_h > TILE_LOG && !tmp)) return P3R_ERR_OOM;
            jobs.push_back({src, dst, d.log_h, 4, false, rot, coef, tmp});
            mats.push_back({dst, logN, 4});
        }
        lmax = std::max(lmax, logN);
    }
    TRY(join_streams(ctx));
    }
    TRY(coset_lde_batch<F>(ctx, jobs, lb));
    uint32_t* dg = arena_alloc<uint32_t>(ctx, tree_digest_words(lmax));
    if (!dg) return P3R_ERR_OOM;
    TRY(commit_tree<F>(ctx, mats, &s->quot_tree, dg));
    TRY(read_cap(ctx, s->quot_tree, cap_out));
    s->phase = PH_QUOT;
    return P3R_OK;
}

// Opened-value layout per instance: see include/p3r.h p3r_open.
template <class F>
static int open_impl(p3r_session* s, const uint32_t zeta_w[4], uint32_t* opened_out, size_t cap_words, size_t* n_words) {
    p3r_ctx* ctx = s->ctx;
    if (s->phase != PH_QUOT) {
        set_err(ctx, "open: wrong phase");
        return P3R_ERR_STATE;
    }
    const p3r_prep* pp = s->prep;
    const uint32_t wnr = ctx->w_m;
    size_t n_inst = pp->inst.size();
    Ext4 zeta;
    std::memcpy(zeta.c, zeta_w, 16);
    s->zeta = zeta;
    // host-side small field helpers for the points
    std::vector<WeightJob> wjobs;
    std::vector<DotJob> djobs;
    uint32_t w_off = 0, o_off = 0, max_n_log = 0, max_w = 1;
    s->open_rounds.assign(4, {});
    std::vector<std::vector<p3r_session::OpenRef>> refs(4);
    for (size_t i = 0; i < n_inst; i++) {
        const InstDev& d = pp->inst[i];
        uint32_t n = 1u << d.log_h;
        max_n_log = std::max(max_n_log, d.log_h);
        // g_n = w_n ; zeta*g
        uint32_t gn = fpow<F>(ctx->gen_m, ((uint64_t)F::P - 1) >> d.log_h);
        Ext4 znext = emul_base<F>(zeta, gn);
        const uint32_t n_inv_m = finv<F>(to_monty<F>(n));
        auto wjob = [&](const Ext4& u, uint32_t off) {  // scale = (u^n - 1) / n
            Ext4 un = u;
            for (uint32_t k = 0; k < d.log_h; k++) un = emul<F>(un, un, wnr);
            return WeightJob{u, emul_base<F>(esub_base<F>(un, F::R), n_inv_m), d.log_h, off};
        };
        uint32_t wz = w_off;
        wjobs.push_back(wjob(zeta, wz));
        w_off += n;
        uint32_t wzn = w_off;
        wjobs.push_back(wjob(znext, wzn));
        w_off += n;
        auto add = [&](const uint32_t* mat, uint32_t width, uint32_t woff) {
            djobs.push_back({mat, d.log_h, width, woff, o_off});
            uint32_t o = o_off;
            o_off += width;
            max_w = std::max(max_w, width);
            return o;
        };
        p3r_session::OpenRef m{(uint32_t)i, 0, 0, {0, 0}, 1, d.main_w};
        m.off[0] = add(s->trace[i], d.main_w, wz);
        if (d.uses_next) {
            m.off[1] = add(s->trace[i], d.main_w, wzn);
            m.n_points = 2;
        }
        refs[0].push_back(m);
        if (d.prep_w) {
            p3r_session::OpenRef p{(uint32_t)i, 2, 0, {0, 0}, 2, d.prep_w};
            p.off[0] = add(d.prep_trace, d.prep_w, wz);
            p.off[1] = add(d.prep_trace, d.prep_w, wzn);
            refs[2].push_back(p);
        }
        if (!d.lookups.empty()) {
            uint32_t pw = d.aux_w() * 4;
            p3r_session::OpenRef p{(uint32_t)i, 3, 0, {0, 0}, 2, pw};
            p.off[0] = add(s->perm[i], pw, wz);
            p.off[1] = add(s->perm[i], pw, wzn);
            refs[3].push_back(p);
        }
        // quotient chunk c: source evaluations on shift_c * H_n, shift_c = GENERATOR * w_NQ^c => u = zeta / shift_c
        uint32_t lq = d.log_h + d.log_qc;
        uint32_t wq = fpow<F>(ctx->gen_m, ((uint64_t)F::P - 1) >> lq);
        for (uint32_t c = 0; c < (1u << d.log_qc); c++) {
            uint32_t shift = fmul<F>(ctx->gen_m, fpow<F>(wq, c));
            Ext4 u = emul_base<F>(zeta, finv<F>(shift));
            uint32_t wc = w_off;
            wjobs.push_back(wjob(u, wc));
            w_off += n;
            p3r_session::OpenRef q{(uint32_t)i, 1, c, {0, 0}, 1, 4};
            q.off[0] = add(s->chunks[i] + (size_t)c * 4 * n, 4, wc);
            refs[1].push_back(q);
        }
    }
    s->open_rounds = refs;
    s->n_opened = o_off;
    Ext4* d_w = arena_alloc<Ext4>(ctx, w_off);
    s->d_opened = arena_alloc<Ext4>(ctx, o_off);
    uint32_t max_chunks = ((1u << max_n_log) + DOT_ROWS - 1) / DOT_ROWS;
    Ext4* partial = arena_alloc<Ext4>(ctx, djobs.size() * (size_t)max_chunks * max_w);
    std::vector<DotTile> tiles;
    for (uint32_t ji = 0; ji < djobs.size(); ji++) {
        const uint32_t chunks = ((1u << djobs[ji].log_n) + DOT_ROWS - 1) / DOT_ROWS;
        for (uint32_t c0 = 0; c0 < djobs[ji].width; c0 += DOT_COLS)
            for (uint32_t ch = 0; ch < chunks; ch++) tiles.push_back({ji, ch, c0});
    }
    WeightJob* d_wj = upload_vec(ctx, wjobs);
    DotJob* d_dj = upload_vec(ctx, djobs);
    DotTile* d_tiles = upload_vec(ctx, tiles);
    if (!d_w || !s->d_opened || !partial || !d_wj || !d_dj || !d_tiles) return P3R_ERR_OOM;
    KT kt_open(ctx, KC_OPEN);
    {
        const uint32_t per_job = std::max(1u, (1u << max_n_log) / 4);  // four weights per thread
        dim3 grid((per_job + 255) / 256, (unsigned)wjobs.size());
        k_bary_weights<F><<<grid, 256, 0, ctx->stream>>>(d_wj, d_w, ctx->tw, ctx->logT, wnr);
        LAUNCH_CHECK();
    }
    {
        k_bary_dot<F><<<(unsigned)tiles.size(), 256, 0, ctx->stream>>>(d_dj, d_tiles, d_w, partial, max_chunks, max_w);
        LAUNCH_CHECK();
        dim3 g2((max_w + 127) / 128, (unsigned)djobs.size());
        k_bary_reduce<F><<<g2, 128, 0, ctx->stream>>>(d_dj, partial, s->d_opened, max_chunks, max_w);
        LAUNCH_CHECK();
    }
    // download; re-order into the per-instance ABI layout
    std::vector<Ext4> host(o_off);
    CUDA_TRY(d2h_async(ctx, host.data(), s->d_opened, (size_t)o_off * 16));
    CUDA_TRY(ctx_wait(ctx));
    std::vector<uint32_t> outw;
    outw.reserve((size_t)o_off * 4);
    auto push = [&](uint32_t off, uint32_t width) {
        for (uint32_t k = 0; k < width; k++)
            for (int c = 0; c < 4; c++) outw.push_back(host[off + k].c[c]);
    };
    for (size_t i = 0; i < n_inst; i++) {
        for (int kind : {0, 2, 3, 1})  // main, prep, perm, quotient chunks
            for (auto& r : refs[kind])
                if (r.inst == i)
                    for (uint32_t p = 0; p < r.n_points; p++) push(r.off[p], r.width);
    }
    s->opened_host = outw;
    *n_words = outw.size();
    if (outw.size() > cap_words) {
        set_err(ctx, "open: output buffer too small");
        return P3R_ERR_BUFFER;
    }
    std::memcpy(opened_out, outw.data(), outw.size() * 4);
    s->phase = PH_OPEN;
    return P3R_OK;
}

static std::vector<uint32_t> arity_schedule(const p3r_fri_params& fri, const std::vector<uint32_t>& heights_desc, bool* ok) {
    std::vector<uint32_t> sched;
    uint32_t log_final = fri.log_blowup + fri.log_final_poly_len;
    uint32_t h = heights_desc[0];
    size_t next = 1;
    *ok = true;
    while (h > log_final) {
        uint32_t k = std::min(fri.max_log_arity, h - log_final);
        if (next < heights_desc.size()) k = std::min(k, h - heights_desc[next]);
        if (k == 0) {
            *ok = false;
            return sched;
        }
        h -= k;
        if (next < heights_desc.size() && heights_desc[next] == h) next++;
        sched.push_back(k);
    }
    if (next != heights_desc.size()) *ok = false;
    return sched;
}

template <class F>
static int fri_begin_impl(p3r_session* s, const uint32_t alpha_w[4], uint32_t* n_rounds_out, uint32_t* log_arities_out) {
    p3r_ctx* ctx = s->ctx;
    if (s->phase != PH_OPEN) {
        set_err(ctx, "fri_begin: wrong phase");
        return P3R_ERR_STATE;
    }
    const p3r_prep* pp = s->prep;
    const uint32_t lb = ctx->fri.log_blowup, wnr = ctx->w_m;
    Ext4 alpha;
    std::memcpy(alpha.c, alpha_w, 16);
    // group (round, matrix, point) by LDE height in commit order; running alpha exponent per height
    struct HGroup {
        std::vector<RoMat> mats;
        uint32_t alpha_count = 0;
        uint32_t log_n = 0;
    };
    std::map<uint32_t, HGroup, std::greater<uint32_t>> groups;
    uint32_t max_exp = 1;
    for (int kind : {0, 1, 2, 3}) {  // rounds in commit order [main, quotient, preprocessed, permutation]
        for (auto& r : s->open_rounds[kind]) {
            const InstDev& d = pp->inst[r.inst];
            uint32_t lh = d.log_h + lb;
            HGroup& g = groups[lh];
            g.log_n = d.log_h;
            RoMat m{};
            size_t N = (size_t)1 << lh;
            switch (kind) {
                case 0: m.lde = s->main_lde[r.inst]; break;
                case 1: m.lde = s->chunk_lde[r.inst] + (size_t)r.chunk * 4 * N; break;
                case 2: m.lde = d.prep_lde; break;
                default: m.lde = s->perm_lde[r.inst]; break;
            }
            m.width = r.width;
            m.n_points = r.n_points;
            for (uint32_t p = 0; p < r.n_points; p++) {
                m.opened_off[p] = r.off[p];
                m.alpha_off[p] = g.alpha_count;
                g.alpha_count += r.width;
            }
            g.mats.push_back(m);
            max_exp = std::max(max_exp, g.alpha_count + 1);
        }
    }
    Ext4* apow = arena_alloc<Ext4>(ctx, max_exp);
    if (!apow) return P3R_ERR_OOM;
    k_ext_powers_asc<F><<<((max_exp + 7) / 8 + 127) / 128, 128, 0, ctx->stream>>>(apow, max_exp, alpha, wnr);  // 8 powers per thread
    LAUNCH_CHECK();
    s->heights.clear();
    s->ro.clear();
    // all heights: one k_ro_prepare launch over every (matrix, point), one k_reduced_openings launch over every row
    std::vector<RoMat> all_mats;
    for (auto& kv : groups) all_mats.insert(all_mats.end(), kv.second.mats.begin(), kv.second.mats.end());
    RoMat* d_all = upload_vec(ctx, all_mats);
    Ext4* coef_all = arena_alloc<Ext4>(ctx, all_mats.size() * 2);
    if (!d_all || !coef_all) return P3R_ERR_OOM;
    {
        const uint32_t nm = (uint32_t)all_mats.size();
        k_ro_prepare<F><<<(nm * 2 * 32 + 127) / 128, 128, 0, ctx->stream>>>(d_all, nm, s->d_opened, apow, coef_all, wnr);
        LAUNCH_CHECK();
    }
    std::vector<RoArgs> ro_jobs;
    uint32_t cta = 0, widest = 1;
    uint64_t ro_bytes = 0;
    size_t mat_off = 0;
    for (auto& kv : groups) {
        uint32_t lh = kv.first;
        HGroup& g = kv.second;
        s->heights.push_back(lh);
        Ext4* ro = arena_alloc<Ext4>(ctx, (size_t)1 << lh);
        if (!ro) return P3R_ERR_OOM;
        RoArgs ra{};
        ra.mats = d_all + mat_off;
        ra.n_mats = (uint32_t)g.mats.size();
        ra.log_h = lh;
        ra.apow = apow;
        ra.coef = coef_all + 2 * mat_off;
        ra.z[0] = s->zeta;
        ra.z[1] = emul_base<F>(s->zeta, fpow<F>(ctx->gen_m, ((uint64_t)F::P - 1) >> g.log_n));
        ra.gen = ctx->gen_m;
        ra.tw = ctx->tw;
        ra.logT = ctx->logT;
        ra.ro = ro;
        ra.wnr = wnr;
        uint64_t wsum = 0;
        for (auto& m : g.mats) {
            wsum += m.width;
            ra.max_width = std::max(ra.max_width, m.width);
        }
        widest = std::max(widest, ra.max_width);
        ra.cta_begin = cta;
        cta += ((1u << lh) + 127) / 128;
        ro_bytes += ((uint64_t)1 << lh) * (4 * wsum + 16);
        mat_off += g.mats.size();
        ro_jobs.push_back(ra);
        s->ro[lh] = ro;
    }
    {
        const RoArgs* d_jobs = upload_vec(ctx, ro_jobs);
        if (!d_jobs) return P3R_ERR_OOM;
        KT kt(ctx, KC_REDUCE, ro_bytes);
        k_reduced_openings<F><<<cta, 128, (size_t)widest * sizeof(Ext4), ctx->stream>>>(d_jobs, (uint32_t)ro_jobs.size());
        LAUNCH_CHECK_C(KC_REDUCE);
    }
    bool ok = false;
    std::vector<uint32_t> sched = arity_schedule(ctx->fri, s->heights, &ok);
    if (!ok || sched.empty() || sched.size() > 32) {
        set_err(ctx, "fri_begin: table heights incompatible with log_final_poly_len/log_blowup");
        return P3R_ERR_INVALID_ARG;
    }
    s->log_max = s->heights[0];
    s->rounds.clear();
    uint32_t h = s->log_max;
    for (size_t r = 0; r < sched.size(); r++) {
        FriRound fr;
        fr.log_arity = sched[r];
        fr.log_len = h;
        fr.vec = (r == 0) ? s->ro[h] : arena_alloc<Ext4>(ctx, (size_t)1 << h);
        if (!fr.vec) return P3R_ERR_OOM;
        h -= sched[r];
        s->rounds.push_back(fr);
        log_arities_out[r] = sched[r];
    }
    s->final_vec = arena_alloc<Ext4>(ctx, (size_t)1 << h);
    if (!s->final_vec) return P3R_ERR_OOM;
    *n_rounds_out = (uint32_t)sched.size();
    s->phase = PH_FRI;
    return P3R_OK;
}

template <class F>
static int fri_commit_impl(p3r_session* s, uint32_t round, uint32_t* cap_out, bool read_back = true) {
    p3r_ctx* ctx = s->ctx;
    if (s->phase != PH_FRI || round >= s->rounds.size()) {
        set_err(ctx, "fri_commit: wrong phase/round");
        return P3R_ERR_STATE;
    }
    FriRound& fr = s->rounds[round];
    uint32_t log_rows = fr.log_len - fr.log_arity;
    if (log_rows < ctx->fri.cap_height) {
        set_err(ctx, "fri_commit: folded height below the Merkle cap");
        return P3R_ERR_INVALID_ARG;
    }
    uint32_t* dg = arena_alloc<uint32_t>(ctx, tree_digest_words(log_rows));
    if (!dg) return P3R_ERR_OOM;
    fr.tree.log_max_h = log_rows;
    fr.tree.digests = dg;
    TRY(build_tree<F>(ctx, reinterpret_cast<const uint32_t*>(fr.vec), 4u << fr.log_arity, log_rows, dg,
                      [](uint32_t) { return (const uint32_t*)nullptr; }));
    return read_back ? read_cap(ctx, fr.tree, cap_out) : P3R_OK;
}

template <class F>
static int fri_fold_impl(p3r_session* s, uint32_t round, const uint32_t beta_w[4], const Ext4* beta_dev = nullptr) {
    p3r_ctx* ctx = s->ctx;
    if (s->phase != PH_FRI || round >= s->rounds.size()) {
        set_err(ctx, "fri_fold: wrong phase/round");
        return P3R_ERR_STATE;
    }
    FriRound& fr = s->rounds[round];
    Ext4 beta{};
    if (beta_w) std::memcpy(beta.c, beta_w, 16);
    uint32_t out_log = fr.log_len - fr.log_arity;
    Ext4* out = (round + 1 < s->rounds.size()) ? s->rounds[round + 1].vec : s->final_vec;
    const Ext4* roll = nullptr;
    auto it = s->ro.find(out_log);
    if (it != s->ro.end() && out_log != s->log_max) roll = it->second;
    uint32_t n_out = 1u << out_log;
    KT kt(ctx, KC_FOLD);
    k_fri_fold<F><<<(n_out + 127) / 128, 128, 0, ctx->stream>>>(fr.vec, out, fr.log_len, fr.log_arity, beta, beta_dev, roll,
                                                                ctx->inv2_m, ctx->tw, ctx->logT, ctx->w_m);
    ctx->kstats.bytes[KC_FOLD] += 16ull * (((size_t)1 << fr.log_len) + n_out);
    LAUNCH_CHECK_C(KC_FOLD);
    return P3R_OK;
}

// Enqueues the final-polynomial kernel and the copy of its coefficients to `coeffs_out`; the caller waits (ctx_wait).
template <class F>
static int fri_final_poly_enqueue(p3r_session* s, uint32_t* coeffs_out) {
    p3r_ctx* ctx = s->ctx;
    if (s->phase != PH_FRI) {
        set_err(ctx, "fri_final_poly: wrong phase");
        return P3R_ERR_STATE;
    }
    uint32_t lf = ctx->fri.log_final_poly_len, n = 1u << lf;
    Ext4* d_c = arena_alloc<Ext4>(ctx, n);
    if (!d_c) return P3R_ERR_OOM;
    k_final_poly<F><<<(n + 63) / 64, 64, 0, ctx->stream>>>(s->final_vec, d_c, lf, ctx->tw, ctx->logT);
    LAUNCH_CHECK();
    CUDA_TRY(d2h_async(ctx, coeffs_out, d_c, (size_t)n * 16));
    return P3R_OK;
}
template <class F>
static int fri_final_poly_impl(p3r_session* s, uint32_t* coeffs_out) {
    p3r_ctx* ctx = s->ctx;
    TRY(fri_final_poly_enqueue<F>(s, coeffs_out));
    CUDA_TRY(ctx_wait(ctx));
    return P3R_OK;
}

static int fri_query_impl(p3r_session* s, const uint32_t* indices, uint32_t nq, uint32_t* out, size_t cap_words,
                          size_t* n_words) {
    p3r_ctx* ctx = s->ctx;
    if (s->phase != PH_FRI) {
        set_err(ctx, "fri_query: wrong phase");
        return P3R_ERR_STATE;
    }
    const uint32_t cap = ctx->fri.cap_height;
    std::vector<GatherSeg> segs;
    uint32_t off = 0;
    std::vector<const Tree*> trees = {&s->main_tree, &s->quot_tree};
    if (s->prep->has_prep) trees.push_back(&s->prep->prep_tree);
    if (s->prep->has_perm) trees.push_back(&s->perm_tree);
    for (const Tree* t : trees) {
        uint32_t tshift = s->log_max - t->log_max_h;
        for (auto& m : t->mats) {
            segs.push_back({0, m.d, m.log_h, m.w, tshift + (t->log_max_h - m.log_h), off});
            off += m.w;
        }
        uint32_t depth = t->log_max_h - cap;
        segs.push_back({1, t->digests, t->log_max_h, depth, tshift, off});
        off += depth * 8;
    }
    uint32_t consumed = 0;
    for (auto& fr : s->rounds) {
        uint32_t arity = 1u << fr.log_arity;
        segs.push_back({2, reinterpret_cast<const uint32_t*>(fr.vec), fr.log_len, fr.log_arity, consumed, off});
        off += (arity - 1) * 4;
        uint32_t log_rows = fr.log_len - fr.log_arity;
        uint32_t depth = log_rows - cap;
        segs.push_back({1, fr.tree.digests, log_rows, depth, consumed + fr.log_arity, off});
        off += depth * 8;
        consumed += fr.log_arity;
    }
    size_t total = (size_t)off * nq;
    *n_words = total;
    if (total > cap_words) {
        set_err(ctx, "fri_query: output buffer too small");
        return P3R_ERR_BUFFER;
    }
    GatherSeg* d_segs = upload_vec(ctx, segs);
    uint32_t* d_idx = (uint32_t*)upload_small(ctx, indices, (size_t)nq * 4);
    uint32_t* d_out = arena_alloc<uint32_t>(ctx, total);
    if (!d_segs || !d_idx || !d_out) return P3R_ERR_OOM;
    dim3 grid(nq, 8);
    k_query_gather<<<grid, 256, 0, ctx->stream>>>(d_segs, (uint32_t)segs.size(), d_idx, off, d_out);
    LAUNCH_CHECK();
    CUDA_TRY(d2h_async(ctx, out, d_out, total * 4));
    CUDA_TRY(ctx_wait(ctx));
    return P3R_OK;
}

template <class F>
static int grind_impl(p3r_ctx* ctx, const uint32_t state[16], const uint32_t* pending, uint32_t n_pending, uint32_t bits,
                      uint32_t* witness_out) {
    if (bits == 0) {
        *witness_out = 0;
        return P3R_OK;
    }
    if (n_pending >= 8 || bits > 30) {
        set_err(ctx, "grind: bad arguments");
        return P3R_ERR_INVALID_ARG;
    }
    uint32_t host[16 + 8 + 1];
    std::memcpy(host, state, 64);
    std::memset(host + 16, 0, 32);
    if (n_pending) std::memcpy(host + 16, pending, (size_t)n_pending * 4);
    host[24] = 0xffffffffu;
    // dedicated staging (never the session's ring: earlier async copies / descriptors of the session may still be in use)
    if (!ctx->grind_pin) {
        CUDA_TRY(cudaHostAlloc((void**)&ctx->grind_pin, 128, cudaHostAllocDefault));
        CUDA_TRY(cudaMalloc((void**)&ctx->grind_dev, 128));
    }
    CUDA_TRY(ctx_wait(ctx));   // a previous grind's upload must have been consumed before its pinned source is rewritten
    std::memcpy(ctx->grind_pin, host, sizeof host);
    uint32_t* d = ctx->grind_dev;
    CUDA_TRY(cudaMemcpyAsync(d, ctx->grind_pin, sizeof host, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t batch = std::max(1u << 16, 4u << bits);
    for (uint64_t base = 0; base < F::P; base += batch) {
        k_grind<F><<<(batch + 127) / 128, 128, 0, ctx->stream>>>(d, d + 16, n_pending, bits, (uint32_t)base, batch, d + 24);
        LAUNCH_CHECK();
        uint32_t best;
        CUDA_TRY(d2h_async(ctx, &best, d + 24, 4));
        CUDA_TRY(ctx_wait(ctx));
        if (best != 0xffffffffu) {
            *witness_out = to_monty<F>(best);
            return P3R_OK;
        }
    }
    set_err(ctx, "grind: no witness");
    return P3R_ERR_POW;
}

// ------------------------------------------------------------------------------------------------
// One-shot prove: host transcript (SURVEY.md A1/A2/A7 order) over the phases. Blob layout: DESIGN.md "Proof blob".
// ------------------------------------------------------------------------------------------------
struct PhaseTimer {   // CUDA events on the session stream, taken from a pool owned by the context (no create/destroy per proof)
    p3r_ctx* ctx;
    size_t used = 0;
    std::vector<std::string> names;
    explicit PhaseTimer(p3r_ctx* c) : ctx(c) { mark("start"); }
    void mark(const char* name) {
        if (used == ctx->phase_ev.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ctx->phase_ev.push_back(e);
        }
        cudaEventRecord(ctx->phase_ev[used++], ctx->stream);
        names.push_back(name);
    }
    void finish() {
        cudaEventSynchronize(ctx->phase_ev[used - 1]);
        ctx->phase_times.clear();
        for (size_t i = 1; i < used; i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ctx->phase_ev[i - 1], ctx->phase_ev[i]);
            ctx->phase_times.push_back({names[i], ms});
        }
    }
};

template <class F>
static int challenger_grind(p3r_ctx* ctx, HostChallenger<F>& ch, uint32_t bits, uint32_t* witness) {
    if (bits == 0) {
        *witness = 0;
        return P3R_OK;
    }
    TRY(grind_impl<F>(ctx, ch.st, ch.in, (uint32_t)ch.n_in, bits, witness));
    ch.observe(*witness);
    if (ch.sample_bits(bits) != 0) {
        set_err(ctx, "grind: device witness rejected by host transcript");
        return P3R_ERR_POW;
    }
    return P3R_OK;
}

template <class F>
static int prove_impl(p3r_ctx* ctx, const p3r_prep* prep, const p3r_matrix_u32* traces, const uint32_t* const* public_values,
                      uint32_t* proof_out, size_t cap_words, size_t* n_words, const p3r_traces* resident = nullptr,
                      const p3r_table_ops* tops = nullptr) {
    p3r_session* s = nullptr;
    TRY(prove_begin_impl<F>(ctx, prep, traces, public_values, &s, resident, tops));
    struct Guard {
        p3r_session* s;
        ~Guard() { delete s; }
    } guard{s};
    PhaseTimer pt(ctx);
    const size_t n_inst = prep->inst.size();
    const size_t capw = (size_t)8 << ctx->fri.cap_height;
    std::vector<uint32_t> blob;
    auto put = [&](const uint32_t* p, size_t n) { blob.insert(blob.end(), p, p + n); };
    HostChallenger<F> ch(&ctx->p2, ctx->host_permute);

    std::vector<uint32_t> main_cap(capw), perm_cap(capw), quot_cap(capw);
    if (ctx->uni_stark && (n_inst != 1 || prep->has_perm || !prep->inst[0].uses_next)) {
        set_err(ctx, "uni-stark mode: exactly one table, no lookups, main trace opened at zeta and zeta*g");
        return P3R_ERR_INVALID_ARG;
    }
    TRY(commit_main_impl<F>(s, main_cap.data()));
    pt.mark("commit_main");
    if (ctx->uni_stark) {
        // p3_uni_stark::prove's transcript head (restated in-tree: recursion/src/types/challenges.rs:44-54,100-140)
        const InstDev& d = prep->inst[0];
        ch.observe(to_monty<F>(d.log_h));
        ch.observe(to_monty<F>(d.log_h));
        ch.observe(to_monty<F>(d.prep_w));
        ch.observe_words(main_cap.data(), capw);
        if (prep->has_prep) ch.observe_words(prep->prep_cap.data(), capw);
        if (d.n_pub) ch.observe_words(public_values[0], d.n_pub);
    } else {
        ch.observe_lifted((uint32_t)n_inst);
        for (auto& d : prep->inst) {
            ch.observe_lifted(d.log_h);
            ch.observe_lifted(d.log_h);
            ch.observe_lifted(d.main_w);
            ch.observe_lifted(1u << d.log_qc);
        }
        ch.observe_words(main_cap.data(), capw);
        for (size_t i = 0; i < n_inst; i++)
            if (prep->inst[i].n_pub) ch.observe_words(public_values[i], prep->inst[i].n_pub);
        for (auto& d : prep->inst) ch.observe_lifted(d.prep_w);
        if (prep->has_prep) ch.observe_words(prep->prep_cap.data(), capw);
    }

    size_t n_perm_inst = 0;
    for (auto& d : prep->inst) n_perm_inst += !d.lookups.empty();
    std::vector<uint32_t> terminals(4 * std::max<size_t>(n_perm_inst, 1));
    uint32_t pa[4] = {0, 0, 0, 0}, pb[4] = {0, 0, 0, 0};
    if (prep->has_perm) {
        ch.sample_ext(pa);
        ch.sample_ext(pb);
    }
    TRY(commit_perm_impl<F>(s, pa, pb, perm_cap.data(), terminals.data()));
    pt.mark("commit_perm");
    if (prep->has_perm) {
        ch.observe_words(perm_cap.data(), capw);
        ch.observe_words(terminals.data(), 4 * n_perm_inst);
    }
    uint32_t alpha[4];
    ch.sample_ext(alpha);
    TRY(commit_quotient_impl<F>(s, alpha, quot_cap.data()));
    pt.mark("commit_quotient");
    ch.observe_words(quot_cap.data(), capw);
    uint32_t zeta[4];
    ch.sample_ext(zeta);

    size_t n_open = 0;
    {
        size_t need = 0;
        for (auto& d : prep->inst)
            need += 4 * ((size_t)d.main_w * 2 + (size_t)d.prep_w * 2 + (size_t)d.aux_w() * 8 + ((size_t)4 << d.log_qc));
        std::vector<uint32_t> tmp(need);
        TRY(open_impl<F>(s, zeta, tmp.data(), tmp.size(), &n_open));
    }
    pt.mark("open");
    // observe in round order [main, quotient, preprocessed, permutation] (recursion/src/generation.rs:474-485)
    {
        std::vector<Ext4> dummy;
        // opened_host is per instance; rebuild round order from the refs (offsets into the device order = download order)
        // Simpler: walk the per-instance blob with a cursor table.
        std::vector<size_t> base(n_inst);
        size_t cur = 0;
        for (size_t i = 0; i < n_inst; i++) {
            base[i] = cur;
            const InstDev& d = prep->inst[i];
            cur += 4 * ((size_t)d.main_w * (1 + (d.uses_next ? 1 : 0)) + (size_t)d.prep_w * 2 + (size_t)d.aux_w() * 8 +
                        ((size_t)4 << d.log_qc));
        }
        const uint32_t* ov = s->opened_host.data();
        for (size_t i = 0; i < n_inst; i++) {  // main: local (+ next)
            const InstDev& d = prep->inst[i];
            ch.observe_words(ov + base[i], 4 * (size_t)d.main_w * (1 + (d.uses_next ? 1 : 0)));
        }
        for (size_t i = 0; i < n_inst; i++) {  // quotient chunks
            const InstDev& d = prep->inst[i];
            size_t off = base[i] + 4 * ((size_t)d.main_w * (1 + (d.uses_next ? 1 : 0)) + (size_t)d.prep_w * 2 + (size_t)d.aux_w() * 8);
            ch.observe_words(ov + off, 4 * ((size_t)4 << d.log_qc));
        }
        for (size_t i = 0; i < n_inst; i++) {  // preprocessed
            const InstDev& d = prep->inst[i];
            if (!d.prep_w) continue;
            size_t off = base[i] + 4 * ((size_t)d.main_w * (1 + (d.uses_next ? 1 : 0)));
            ch.observe_words(ov + off, 4 * (size_t)d.prep_w * 2);
        }
        for (size_t i = 0; i < n_inst; i++) {  // permutation
            const InstDev& d = prep->inst[i];
            if (d.lookups.empty()) continue;
            size_t off = base[i] + 4 * ((size_t)d.main_w * (1 + (d.uses_next ? 1 : 0)) + (size_t)d.prep_w * 2);
            ch.observe_words(ov + off, 4 * (size_t)d.aux_w() * 8);
        }
    }
    uint32_t alpha_fri[4];
    ch.sample_ext(alpha_fri);
    uint32_t n_rounds = 0, log_arities[32];
    TRY(fri_begin_impl<F>(s, alpha_fri, &n_rounds, log_arities));
    pt.mark("fri_reduce");
    std::vector<uint32_t> fri_caps(n_rounds * capw), commit_pow(n_rounds);
    std::vector<uint32_t> final_poly((size_t)4 << ctx->fri.log_final_poly_len);
    bool final_poly_done = false;
    if (ctx->fri.commit_pow_bits == 0 && ctx->dev_fri_transcript && n_rounds > 0) {
        // No commit-phase PoW: the rounds' caps are observed and the betas sampled by a one-warp kernel, so all rounds are
        // enqueued without a host round trip; afterwards the host challenger replays the same operations on the caps.
        DevChallenger hc{};
        std::memcpy(hc.st, ch.st, sizeof hc.st);
        std::memcpy(hc.in, ch.in, sizeof hc.in);
        std::memcpy(hc.out, ch.out, sizeof hc.out);
        hc.n_in = (uint32_t)ch.n_in;
        hc.n_out = (uint32_t)ch.n_out;
        DevChallenger* d_ch = reinterpret_cast<DevChallenger*>(upload_small(ctx, &hc, sizeof hc, /*device_mutable=*/true));
        Ext4* d_beta = arena_alloc<Ext4>(ctx, n_rounds);
        if (!d_ch || !d_beta) return P3R_ERR_OOM;
        for (uint32_t r = 0; r < n_rounds; r++) {
            TRY(fri_commit_impl<F>(s, r, nullptr, false));
            const Tree& t = s->rounds[r].tree;
            k_fri_round_transcript<F><<<1, 32, 0, ctx->stream>>>(d_ch, t.digests + t.level_off(t.log_max_h - ctx->fri.cap_height) * 8,
                                                                 (uint32_t)capw, d_beta + r, ctx->d_p2);
            LAUNCH_CHECK();
            TRY(fri_fold_impl<F>(s, r, nullptr, d_beta + r));
        }
        for (uint32_t r = 0; r < n_rounds; r++) {
            const Tree& t = s->rounds[r].tree;
            CUDA_TRY(d2h_async(ctx, fri_caps.data() + r * capw, t.digests + t.level_off(t.log_max_h - ctx->fri.cap_height) * 8,
                                     capw * 4));
        }
        std::vector<Ext4> dev_betas(n_rounds);
        CUDA_TRY(d2h_async(ctx, dev_betas.data(), d_beta, n_rounds * sizeof(Ext4)));
        // the final polynomial depends only on the last fold: same wait as the caps and betas (one host round trip less)
        TRY(fri_final_poly_enqueue<F>(s, final_poly.data()));
        final_poly_done = true;
        CUDA_TRY(ctx_wait(ctx));
        for (uint32_t r = 0; r < n_rounds; r++) {
            ch.observe_words(fri_caps.data() + r * capw, capw);
            commit_pow[r] = 0;
            uint32_t beta[4];
            ch.sample_ext(beta);
            if (std::memcmp(beta, dev_betas[r].c, 16) != 0) {
                set_err(ctx, "device FRI transcript diverged from the host challenger");
                return P3R_ERR_CUDA;
            }
        }
    } else {
        for (uint32_t r = 0; r < n_rounds; r++) {
            TRY(fri_commit_impl<F>(s, r, fri_caps.data() + r * capw));
            ch.observe_words(fri_caps.data() + r * capw, capw);
            TRY(challenger_grind<F>(ctx, ch, ctx->fri.commit_pow_bits, &commit_pow[r]));
            uint32_t beta[4];
            ch.sample_ext(beta);
            TRY(fri_fold_impl<F>(s, r, beta));
        }
    }
    if (!final_poly_done) TRY(fri_final_poly_impl<F>(s, final_poly.data()));
    pt.mark("fri_commit_phase");
    ch.observe_words(final_poly.data(), final_poly.size());
    for (uint32_t r = 0; r < n_rounds; r++) ch.observe(to_monty<F>(log_arities[r]));
    uint32_t query_pow = 0;
    TRY(challenger_grind<F>(ctx, ch, ctx->fri.query_pow_bits, &query_pow));
    pt.mark("grind");
    std::vector<uint32_t> indices(ctx->fri.num_queries);
    for (auto& ix : indices) ix = ch.sample_bits(s->log_max);

    uint32_t hdr[5] = {0x50335250u, (uint32_t)n_inst, (uint32_t)prep->has_perm, (uint32_t)prep->has_prep, (uint32_t)capw};
    put(hdr, 5);
    for (auto& d : prep->inst) blob.push_back(d.log_h);
    put(main_cap.data(), capw);
    if (prep->has_perm) put(perm_cap.data(), capw);
    put(quot_cap.data(), capw);
    if (prep->has_perm) put(terminals.data(), 4 * n_perm_inst);
    put(s->opened_host.data(), s->opened_host.size());
    blob.push_back(n_rounds);
    put(log_arities, n_rounds);
    put(fri_caps.data(), fri_caps.size());
    put(commit_pow.data(), n_rounds);
    put(final_poly.data(), final_poly.size());
    blob.push_back(query_pow);
    size_t head = blob.size();
    size_t qwords = 0;
    // query the size first with a zero-capacity call
    {
        int rc = fri_query_impl(s, indices.data(), (uint32_t)indices.size(), nullptr, 0, &qwords);
        if (rc != P3R_ERR_BUFFER && rc != P3R_OK) return rc;
    }
    *n_words = head + qwords;
    if (*n_words > cap_words) {
        set_err(ctx, "prove: proof buffer too small");
        return P3R_ERR_BUFFER;
    }
    std::memcpy(proof_out, blob.data(), head * 4);
    TRY(fri_query_impl(s, indices.data(), (uint32_t)indices.size(), proof_out + head, cap_words - head, &qwords));
    pt.mark("query");
    pt.finish();
    ctx->phase_times.push_back({"host_challenger", (float)ch.host_ms});
    ctx->phase_times.push_back({"host_perms", (float)ch.n_perms});
    kstats_collect(ctx);
    return P3R_OK;
}

// ------------------------------------------------------------------------------------------------
// isolated entry points
// ------------------------------------------------------------------------------------------------
template <class F>
static int coset_lde_host_impl(p3r_ctx* ctx, const p3r_matrix_u32* in, uint32_t log_blowup, uint32_t* out) {
    if (!is_pow2(in->height) || !in->width) {
        set_err(ctx, "coset_lde: height must be a power of two");
        return P3R_ERR_INVALID_ARG;
    }
    ctx->arena.reset();
    ring_reset(ctx);
    uint32_t log_n = ilog2(in->height);
    size_t n = in->height, N = n << log_blowup, w = in->width;
    uint32_t* rm = arena_alloc<uint32_t>(ctx, N * w);
    uint32_t* cm = arena_alloc<uint32_t>(ctx, n * w);
    uint32_t* coef = arena_alloc<uint32_t>(ctx, n * w);
    uint32_t* lde = arena_alloc<uint32_t>(ctx, N * w);
    uint32_t* tmp = log_n > TILE_LOG ? arena_alloc<uint32_t>(ctx, N * w) : nullptr;
    if (!rm || !cm || !coef || !lde || (log_n > TILE_LOG && !tmp)) return P3R_ERR_OOM;
    TRY(upload_matrix(ctx, *in, rm, cm));
    TRY(coset_lde<F>(ctx, cm, lde, log_n, (uint32_t)w, log_blowup, true, 0, coef, tmp));
    dim3 grid((unsigned)((N + 31) / 32), (unsigned)((w + 31) / 32)), block(32, 8);
    k_transpose_out<<<grid, block, 0, ctx->stream>>>(lde, rm, (uint32_t)N, (uint32_t)w);
    LAUNCH_CHECK();
    CUDA_TRY(d2h_async(ctx, out, rm, N * w * 4));
    CUDA_TRY(ctx_wait(ctx));
    return P3R_OK;
}
template <class F>
static int mmcs_commit_host_impl(p3r_ctx* ctx, uint32_t n_mats, const p3r_matrix_u32* mats, uint32_t* cap_out) {
    ctx->arena.reset();
    ring_reset(ctx);
    std::vector<MatRef> refs;
    uint32_t lmax = 0;
    for (uint32_t i = 0; i < n_mats; i++) {
        if (!is_pow2(mats[i].height) || !mats[i].width) {
            set_err(ctx, "mmcs_commit: heights must be powers of two");
            return P3R_ERR_INVALID_ARG;
        }
        size_t words = (size_t)mats[i].height * mats[i].width;
        uint32_t* rm = arena_alloc<uint32_t>(ctx, words);
        uint32_t* cm = arena_alloc<uint32_t>(ctx, words);
        if (!rm || !cm) return P3R_ERR_OOM;
        TRY(upload_matrix(ctx, mats[i], rm, cm));
        refs.push_back({cm, ilog2(mats[i].height), mats[i].width});
        lmax = std::max(lmax, refs.back().log_h);
    }
    uint32_t* dg = arena_alloc<uint32_t>(ctx, tree_digest_words(lmax));
    if (!dg) return P3R_ERR_OOM;
    Tree t;
    TRY(commit_tree<F>(ctx, refs, &t, dg));
    return read_cap(ctx, t, cap_out);
}
template <class F>
static int permute_host_impl(p3r_ctx* ctx, uint32_t* states, uint32_t n) {
    ctx->arena.reset();
    uint32_t* d = arena_alloc<uint32_t>(ctx, (size_t)n * 16);
    if (!d) return P3R_ERR_OOM;
    CUDA_TRY(cudaMemcpyAsync(d, states, (size_t)n * 64, cudaMemcpyHostToDevice, ctx->stream));
    k_permute_states<F><<<(n + 127) / 128, 128, 0, ctx->stream>>>(d, n);
    LAUNCH_CHECK();
    CUDA_TRY(d2h_async(ctx, states, d, (size_t)n * 64));
    CUDA_TRY(ctx_wait(ctx));
    return P3R_OK;
}
template <class F>
static int permute_w_impl(p3r_ctx* ctx, const Poseidon2ConstsW& k, uint32_t* states, uint32_t n) {
    ctx->arena.reset();
    uint32_t* d = arena_alloc<uint32_t>(ctx, (size_t)n * k.width);
    Poseidon2ConstsW* dk = arena_alloc<Poseidon2ConstsW>(ctx, 1);
    if (!d || !dk) return P3R_ERR_OOM;
    CUDA_TRY(cudaMemcpyAsync(dk, &k, sizeof k, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d, states, (size_t)n * k.width * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (k.width == 24) k_permute_states_w<F, 24><<<(n + 127) / 128, 128, 0, ctx->stream>>>(d, n, dk);
    else k_permute_states_w<F, 16><<<(n + 127) / 128, 128, 0, ctx->stream>>>(d, n, dk);
    LAUNCH_CHECK();
    CUDA_TRY(d2h_async(ctx, states, d, (size_t)n * k.width * 4));
    CUDA_TRY(ctx_wait(ctx));   // `k` (pageable source of the first copy) must outlive it
    return P3R_OK;
}
template <class F>
static int run_chains_impl(p3r_ctx* ctx, const p3r_poseidon2_chain_ops* ops, uint32_t* inputs_out, uint32_t* outputs_out) {
    ctx->arena.reset();
    const size_t n = ops->n_rows;
    if (n == 0) return P3R_OK;
    uint8_t* d_flags = arena_alloc<uint8_t>(ctx, 4 * n);
    uint32_t* d_val = arena_alloc<uint32_t>(ctx, n * 16);
    uint32_t* d_io = arena_alloc<uint32_t>(ctx, n * 32);
    if (!d_flags || !d_val || !d_io) return P3R_ERR_OOM;
    CUDA_TRY(cudaMemcpyAsync(d_flags, ops->new_start, n, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_flags + n, ops->merkle_path, n, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_flags + 2 * n, ops->mmcs_bit, n, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_flags + 3 * n, ops->witness_mask, n, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_val, ops->values, n * 64, cudaMemcpyHostToDevice, ctx->stream));
    P2ChainArgs a{d_flags, d_flags + n, d_flags + 2 * n, d_flags + 3 * n, d_val, (uint32_t)n, d_io, d_io + n * 16};
    k_poseidon2_chains<F><<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(a);
    LAUNCH_CHECK_C(KC_MISC);
    CUDA_TRY(d2h_async(ctx, inputs_out, d_io, n * 64));
    CUDA_TRY(d2h_async(ctx, outputs_out, d_io + n * 16, n * 64));
    CUDA_TRY(ctx_wait(ctx));
    return P3R_OK;
}
// Mixed-height version of the commit benchmark: one batched LDE + one MMCS commit over several synthetic matrices.
template <class F>
static int bench_commit_multi_impl(p3r_ctx* ctx, uint32_t n_mats, const uint32_t* log_heights, const uint32_t* widths, uint32_t iters,
                                   uint64_t seed, float* ms_out) {
    ctx->arena.reset();
    ring_reset(ctx);
    const uint32_t lb = ctx->fri.log_blowup;
    std::vector<LdeJob> jobs;
    std::vector<MatRef> mats;
    uint32_t lmax = 0;
    for (uint32_t i = 0; i < n_mats; i++) {
        const size_t n = (size_t)1 << log_heights[i], N = n << lb, w = widths[i];
        uint32_t* cm = arena_alloc<uint32_t>(ctx, n * w);
        uint32_t* coef = arena_alloc<uint32_t>(ctx, n * w);
        uint32_t* lde = arena_alloc<uint32_t>(ctx, N * w);
        uint32_t* tmp = log_heights[i] > TILE_LOG ? arena_alloc<uint32_t>(ctx, N * w) : nullptr;
        if (!cm || !coef || !lde || (log_heights[i] > TILE_LOG && !tmp)) return P3R_ERR_OOM;
        k_fill_random<F><<<(unsigned)((n * w + 255) / 256), 256, 0, ctx->stream>>>(cm, n * w, seed + i);
        LAUNCH_CHECK();
        jobs.push_back({cm, lde, log_heights[i], widths[i], true, 0, coef, tmp});
        mats.push_back({lde, log_heights[i] + lb, widths[i]});
        lmax = std::max(lmax, log_heights[i] + lb);
    }
    uint32_t* dg = arena_alloc<uint32_t>(ctx, tree_digest_words(lmax));
    if (!dg) return P3R_ERR_OOM;
    TRY(coset_lde_batch<F>(ctx, jobs, lb));   // warm-up (tables, attributes)
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventCreate(&e2);
    float t_lde = 0, t_tree = 0;
    for (uint32_t it = 0; it < iters; it++) {
        size_t pin_mark = ctx->pin_used;
        cudaEventRecord(e0, ctx->stream);
        TRY(coset_lde_batch<F>(ctx, jobs, lb));
        cudaEventRecord(e1, ctx->stream);
        Tree t;
        TRY(commit_tree<F>(ctx, mats, &t, dg));
        cudaEventRecord(e2, ctx->stream);
        CUDA_TRY(cudaEventSynchronize(e2));
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, e0, e1);
        cudaEventElapsedTime(&b, e1, e2);
        t_lde += a;
        t_tree += b;
        ctx->pin_used = pin_mark;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    kstats_collect(ctx);
    ms_out[0] = t_lde / iters;
    ms_out[1] = t_tree / iters;
    return P3R_OK;
}
template <class F>
static int bench_commit_impl(p3r_ctx* ctx, uint32_t log_height, uint32_t width, uint32_t iters, uint64_t seed, float* ms_out) {
    ctx->arena.reset();
    ring_reset(ctx);
    const uint32_t lb = ctx->fri.log_blowup;
    size_t n = (size_t)1 << log_height, N = n << lb, w = width;
    uint32_t* cm = arena_alloc<uint32_t>(ctx, n * w);
    uint32_t* coef = arena_alloc<uint32_t>(ctx, n * w);
    uint32_t* lde = arena_alloc<uint32_t>(ctx, N * w);
    uint32_t* tmp = log_height > TILE_LOG ? arena_alloc<uint32_t>(ctx, N * w) : nullptr;
    uint32_t* dg = arena_alloc<uint32_t>(ctx, tree_digest_words(log_height + lb));
    if (!cm || !coef || !lde || !dg || (log_height > TILE_LOG && !tmp)) return P3R_ERR_OOM;
    k_fill_random<F><<<(unsigned)((n * w + 255) / 256), 256, 0, ctx->stream>>>(cm, n * w, seed);
    LAUNCH_CHECK();
    TRY(coset_lde<F>(ctx, cm, lde, log_height, width, lb, true, 0, coef, tmp));  // warm-up (tables, attributes)
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventCreate(&e2);
    float t_lde = 0, t_tree = 0;
    for (uint32_t it = 0; it < iters; it++) {
        size_t pin_mark = ctx->pin_used;
        cudaEventRecord(e0, ctx->stream);
        TRY(coset_lde<F>(ctx, cm, lde, log_height, width, lb, true, 0, coef, tmp));
        cudaEventRecord(e1, ctx->stream);
        Tree t;
        TRY(commit_tree<F>(ctx, {{lde, log_height + lb, width}}, &t, dg));
        cudaEventRecord(e2, ctx->stream);
        CUDA_TRY(cudaEventSynchronize(e2));
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, e0, e1);
        cudaEventElapsedTime(&b, e1, e2);
        t_lde += a;
        t_tree += b;
        ctx->pin_used = pin_mark;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    kstats_collect(ctx);
    ms_out[0] = t_lde / iters;
    ms_out[1] = t_tree / iters;
    ms_out[2] = 0;
    return P3R_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
#define DISPATCH(ctx, call)                          \
    ((ctx)->field_id == P3R_FIELD_KOALABEAR ? [&] {  \
        using F = KoalaBear;                         \
        return call;                                 \
    }()                                              \
                                            : [&] {  \
                                                  using F = BabyBear; \
                                                  return call;        \
                                              }())

static thread_local std::string g_noctx_err;

// The one-thread-per-permutation kernels read the Poseidon2 constants from `__constant__ c_p2[field]`, which is one copy per
// device. Live contexts of one (device, field) therefore have to agree on the constants: the registry loads the symbol for the
// first context, lets later contexts with the SAME constants share it (no copy while kernels may be in flight) and refuses a
// context whose constants differ (P3R_ERR_UNSUPPORTED) instead of silently changing the hashes of the others.
struct P2Slot {
    int refs = 0;
    Poseidon2Consts k{};
};
static std::mutex g_p2_mu;
static std::map<std::pair<int, int>, P2Slot> g_p2_slots;
static int p2_registry_acquire(int device, int field_id, const Poseidon2Consts& k) {
    std::lock_guard<std::mutex> lock(g_p2_mu);
    P2Slot& slot = g_p2_slots[{device, field_id}];
    if (slot.refs > 0) {
        if (std::memcmp(&slot.k, &k, sizeof k) != 0) {
            g_noctx_err = "ctx_create: a live context on this device uses different Poseidon2 constants for this field";
            return P3R_ERR_UNSUPPORTED;
        }
        slot.refs++;
        return P3R_OK;
    }
    if (cudaMemcpyToSymbol(c_p2, &k, sizeof(Poseidon2Consts), (size_t)field_id * sizeof(Poseidon2Consts)) != cudaSuccess) {
        g_noctx_err = std::string("ctx_create: ") + cudaGetErrorString(cudaGetLastError());
        return P3R_ERR_CUDA;
    }
    slot.k = k;
    slot.refs = 1;
    return P3R_OK;
}
static void p2_registry_release(int device, int field_id) {
    std::lock_guard<std::mutex> lock(g_p2_mu);
    auto it = g_p2_slots.find({device, field_id});
    if (it != g_p2_slots.end() && it->second.refs > 0) it->second.refs--;
}

extern "C" {

uint32_t p3r_abi_version(void) { return 1; }
const char* p3r_build_info(void) { return "libp3r_b200 sm_100a; fields: koala-bear, baby-bear; ext degree 4; poseidon2 width 16"; }

int p3r_ctx_create(int device, const p3r_field_desc* field, const p3r_poseidon2_consts* p2, const p3r_fri_params* fri,
                   p3r_ctx** out) {
    if (!field || !p2 || !fri || !out || !p2->external_rc || !p2->internal_rc || !p2->internal_diag) return P3R_ERR_INVALID_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) {
        g_noctx_err = "no usable CUDA device (this library has no CPU fallback)";
        cudaGetLastError();
        return P3R_ERR_CUDA;
    }
    uint32_t want_p = field->field_id == P3R_FIELD_KOALABEAR ? KoalaBear::P : BabyBear::P;
    if (field->field_id > 1 || field->p != want_p) return P3R_ERR_UNSUPPORTED;
    uint32_t rp = field->field_id == P3R_FIELD_KOALABEAR ? KoalaBear::ROUNDS_P : BabyBear::ROUNDS_P;
    uint32_t sb = field->field_id == P3R_FIELD_KOALABEAR ? KoalaBear::SBOX : BabyBear::SBOX;
    if (p2->width != 16 || p2->rounds_f != 8 || p2->rounds_p != rp || p2->sbox_degree != sb) return P3R_ERR_UNSUPPORTED;
    if (fri->max_log_arity < 1 || fri->max_log_arity > 4 || fri->log_blowup < 1 || fri->log_final_poly_len > 6 ||
        fri->query_pow_bits > 30 || fri->commit_pow_bits > 30)
        return P3R_ERR_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return P3R_ERR_CUDA;
    auto* ctx = new p3r_ctx();
    ctx->device = device;
    ctx->field_id = (int)field->field_id;
    ctx->field = *field;
    ctx->fri = *fri;
    std::memcpy(ctx->p2.ext_rc, p2->external_rc, 8 * 16 * 4);
    std::memset(ctx->p2.int_rc, 0, sizeof ctx->p2.int_rc);
    std::memcpy(ctx->p2.int_rc, p2->internal_rc, rp * 4);
    std::memcpy(ctx->p2.diag, p2->internal_diag, 16 * 4);
    {
        // structured diagonal (shift/add products) only when the caller's diagonal is the p3 one for this field
        bool fast = true;
        for (int i = 0; i < 16; i++) {
            uint32_t want = ctx->field_id == 0 ? to_monty<KoalaBear>(diag_spec_canonical<KoalaBear>(i))
                                               : to_monty<BabyBear>(diag_spec_canonical<BabyBear>(i));
            fast = fast && ctx->p2.diag[i] == want;
        }
        ctx->p2.fast_diag = fast ? 1u : 0u;
    }
    ctx->host_permute = ctx->field_id == 0 ? pick_host_permute<KoalaBear>(ctx->p2) : pick_host_permute<BabyBear>(ctx->p2);
    if (ctx->field_id == 0) {
        ctx->w_m = to_monty<KoalaBear>(field->w);
        ctx->gen_m = to_monty<KoalaBear>(field->generator);
        ctx->inv2_m = finv<KoalaBear>(to_monty<KoalaBear>(2));
    } else {
        ctx->w_m = to_monty<BabyBear>(field->w);
        ctx->gen_m = to_monty<BabyBear>(field->generator);
        ctx->inv2_m = finv<BabyBear>(to_monty<BabyBear>(2));
    }
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) ctx->n_sms = (uint32_t)sms;
    }
    if (const char* e = getenv("P3R_UPLOAD_SKIP")) ctx->skip_equal_uploads = atoi(e) != 0;
    if (const char* e = getenv("P3R_LDE_SMALL_CTA")) ctx->lde_small_cta = atoi(e) != 0;
    if (const char* e = getenv("P3R_LDE_STREAMS")) ctx->lde_streams = (uint32_t)std::max(1, std::min(atoi(e), (int)p3r_ctx::N_AUX));
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
    ctx->pin_size = ctx->dstage_size = (size_t)8 << 20;
    ok = ok && cudaHostAlloc((void**)&ctx->pin, ctx->pin_size, cudaHostAllocDefault) == cudaSuccess;
    ctx->pin_out_size = (size_t)4 << 20;
    ok = ok && cudaHostAlloc((void**)&ctx->pin_out, ctx->pin_out_size, cudaHostAllocDefault) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&ctx->dstage, ctx->dstage_size) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&ctx->d_p2, sizeof(Poseidon2Consts)) == cudaSuccess;
    ok = ok && cudaMemcpy(ctx->d_p2, &ctx->p2, sizeof(Poseidon2Consts), cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) g_noctx_err = std::string("ctx_create: ") + cudaGetErrorString(cudaGetLastError());
    int reg_rc = ok ? p2_registry_acquire(device, ctx->field_id, ctx->p2) : P3R_ERR_CUDA;
    if (reg_rc != P3R_OK) {
        if (ctx->stream) cudaStreamDestroy(ctx->stream);
        if (ctx->pin) cudaFreeHost(ctx->pin);
        if (ctx->pin_out) cudaFreeHost(ctx->pin_out);
        if (ctx->dstage) cudaFree(ctx->dstage);
        if (ctx->d_p2) cudaFree(ctx->d_p2);
        delete ctx;
        return reg_rc;
    }
    *out = ctx;
    return P3R_OK;
}
void p3r_ctx_destroy(p3r_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx_wait(ctx);
    ctx->arena.destroy();
    for (auto& kv : ctx->gtables) {
        cudaFree(kv.second.lo);
        cudaFree(kv.second.hi);
    }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    for (auto e : ctx->phase_ev) cudaEventDestroy(e);
    if (ctx->tws) cudaFree(ctx->tws);
    if (ctx->d_p2) cudaFree(ctx->d_p2);
    if (ctx->d_p2w) cudaFree(ctx->d_p2w);
    if (ctx->pin) cudaFreeHost(ctx->pin);
    if (ctx->pin_out) cudaFreeHost(ctx->pin_out);
    if (ctx->dstage) cudaFree(ctx->dstage);
    if (ctx->grind_pin) cudaFreeHost(ctx->grind_pin);
    if (ctx->grind_dev) cudaFree(ctx->grind_dev);
    for (int i = 0; i < p3r_ctx::N_AUX; i++) {
        if (ctx->aux[i]) cudaStreamDestroy(ctx->aux[i]);
        if (ctx->aux_ev[i]) cudaEventDestroy(ctx->aux_ev[i]);
    }
    if (ctx->aux_ev[p3r_ctx::N_AUX]) cudaEventDestroy(ctx->aux_ev[p3r_ctx::N_AUX]);
    cudaStreamDestroy(ctx->stream);
    p2_registry_release(ctx->device, ctx->field_id);
    delete ctx;
}
int p3r_ctx_set_stream_priority(p3r_ctx* ctx, int high) {
    if (!ctx) return P3R_ERR_INVALID_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(ctx_wait(ctx));
    int least = 0, greatest = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    const int prio = high ? greatest : least;
    if (prio == ctx->stream_prio) return P3R_OK;
    cudaStream_t fresh = nullptr;
    CUDA_TRY(cudaStreamCreateWithPriority(&fresh, cudaStreamNonBlocking, prio));
    cudaStreamDestroy(ctx->stream);
    ctx->stream = fresh;
    ctx->stream_prio = prio;
    for (int i = 0; i < p3r_ctx::N_AUX; i++)       // the side streams are re-created at the next fork with the new priority
        if (ctx->aux[i]) {
            cudaStreamDestroy(ctx->aux[i]);
            ctx->aux[i] = nullptr;
            cudaEventDestroy(ctx->aux_ev[i]);
            ctx->aux_ev[i] = nullptr;
        }
    return P3R_OK;
}
const char* p3r_last_error(const p3r_ctx* ctx) { return ctx ? ctx->err.c_str() : g_noctx_err.c_str(); }

int p3r_prep_commit(p3r_ctx* ctx, uint32_t n_inst, const p3r_instance_desc* descs, const p3r_matrix_u32* prep, p3r_prep** out,
                    uint32_t* cap_out, uint32_t* has_prep_out) {
    if (!ctx || !descs || !out || n_inst == 0) return P3R_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    return DISPATCH(ctx, prep_commit_impl<F>(ctx, n_inst, descs, prep, out, cap_out, has_prep_out));
}
static int load_consts_w(const p3r_ctx* ctx, const p3r_poseidon2_consts* c, Poseidon2ConstsW* out) {
    const uint32_t sb = ctx->field_id == P3R_FIELD_KOALABEAR ? KoalaBear::SBOX : BabyBear::SBOX;
    if (!c || !c->external_rc || !c->internal_rc || !c->internal_diag || (c->width != 16 && c->width != 24) || c->rounds_f != 8 ||
        c->rounds_p == 0 || c->rounds_p > 32 || c->sbox_degree != sb)
        return P3R_ERR_UNSUPPORTED;
    std::memset(out, 0, sizeof *out);
    out->width = c->width;
    out->rounds_p = c->rounds_p;
    std::memcpy(out->ext_rc, c->external_rc, (size_t)8 * c->width * 4);
    std::memcpy(out->int_rc, c->internal_rc, (size_t)c->rounds_p * 4);
    std::memcpy(out->diag, c->internal_diag, (size_t)c->width * 4);
    return P3R_OK;
}
int p3r_ctx_set_leaf_hasher(p3r_ctx* ctx, const p3r_poseidon2_consts* w24) {
    if (!ctx) return P3R_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    CUDA_TRY(ctx_wait(ctx));
    if (!w24) {
        if (ctx->d_p2w) cudaFree(ctx->d_p2w);
        ctx->d_p2w = nullptr;
        return P3R_OK;
    }
    Poseidon2ConstsW k;
    int rc = load_consts_w(ctx, w24, &k);
    if (rc != P3R_OK || k.width != 24) {
        set_err(ctx, "set_leaf_hasher: width-24 Poseidon2 constants of this field expected");
        return P3R_ERR_UNSUPPORTED;
    }
    if (!ctx->d_p2w) CUDA_TRY(cudaMalloc((void**)&ctx->d_p2w, sizeof k));
    CUDA_TRY(cudaMemcpy(ctx->d_p2w, &k, sizeof k, cudaMemcpyHostToDevice));
    return P3R_OK;
}
int p3r_poseidon2_permute_w(p3r_ctx* ctx, const p3r_poseidon2_consts* consts, uint32_t* states, uint32_t n) {
    if (!ctx || !states) return P3R_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    Poseidon2ConstsW k;
    int rc = load_consts_w(ctx, consts, &k);
    if (rc != P3R_OK) {
        set_err(ctx, "poseidon2_permute_w: width 16 or 24 constants of this field expected");
        return rc;
    }
    return DISPATCH(ctx, permute_w_impl<F>(ctx, k, states, n));
}
int p3r_ctx_set_uni_stark(p3r_ctx* ctx, int on) {
    if (!ctx) return P3R_ERR_INVALID_ARG;
    ctx->uni_stark = on != 0;
    return P3R_OK;
}
int p3r_ctx_set_conventions(p3r_ctx* ctx, const p3r_conventions* conv) {
    if (!ctx || !conv || conv->logup_negate > 1 || conv->logup_first_power > 1 || conv->logup_descending > 1) return P3R_ERR_INVALID_ARG;
    ctx->conv = *conv;
    return P3R_OK;
}
void p3r_set_wait_mode(int mode) { g_wait_mode.store(mode < 0 || mode > 3 ? 1 : mode, std::memory_order_relaxed); }
void p3r_prep_free(p3r_prep* prep) {
    if (!prep) return;
    cudaSetDevice(prep->ctx->device);
    ctx_wait(prep->ctx);
    for (void* p : prep->owned) cudaFree(p);
    delete prep;
}
int p3r_prove_begin(p3r_ctx* ctx, const p3r_prep* prep, const p3r_matrix_u32* traces, const uint32_t* const* public_values,
                    p3r_session** out) {
    if (!ctx || !prep || !traces || !out) return P3R_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    return DISPATCH(ctx, prove_begin_impl<F>(ctx, prep, traces, public_values, out));
}
int p3r_commit_main(p3r_session* s, uint32_t* cap_out) {
    if (!s || !cap_out) return P3R_ERR_INVALID_ARG;
    return DISPATCH(s->ctx, commit_main_impl<F>(s, cap_out));
}
int p3r_commit_perm(p3r_session* s, const uint32_t alpha[4], const uint32_t beta[4], uint32_t* cap_out, uint32_t* terminals_out) {
    if (!s) return P3R_ERR_INVALID_ARG;
    return DISPATCH(s->ctx, commit_perm_impl<F>(s, alpha, beta, cap_out, terminals_out));
}
int p3r_commit_quotient(p3r_session* s, const uint32_t alpha[4], uint32_t* cap_out) {
    if (!s || !cap_out) return P3R_ERR_INVALID_ARG;
    return DISPATCH(s->ctx, commit_quotient_impl<F>(s, alpha, cap_out));
}
int p3r_open(p3r_session* s, const uint32_t zeta[4], uint32_t* opened_out, size_t cap_words, size_t* n_words) {
    if (!s || !n_words) return P3R_ERR_INVALID_ARG;
    return DISPATCH(s->ctx, open_impl<F>(s, zeta, opened_out, cap_words, n_words));
}
int p3r_fri_begin(p3r_session* s, const uint32_t alpha_fri[4], uint32_t* n_rounds_out, uint32_t* log_arities_out) {
    if (!s || !n_rounds_out || !log_arities_out) return P3R_ERR_INVALID_ARG;
    return DISPATCH(s->ctx, fri_begin_impl<F>(s, alpha_fri, n_rounds_out, log_arities_out));
}
int p3r_fri_commit(p3r_session* s, uint32_t round, uint32_t* cap_out) {
    if (!s || !cap_out) return P3R_ERR_INVALID_ARG;
    return DISPATCH(s->ctx, fri_commit_impl<F>(s, round, cap_out));
}
int p3r_fri_fold(p3r_session* s, uint32_t round, const uint32_t beta[4]) {
    if (!s) return P3R_ERR_INVALID_ARG;
    return DISPATCH(s->ctx, fri_fold_impl<F>(s, round, beta));
}
int p3r_fri_final_poly(p3r_session* s, uint32_t* coeffs_out) {
    if (!s || !coeffs_out) return P3R_ERR_INVALID_ARG;
    return DISPATCH(s->ctx, fri_final_poly_impl<F>(s, coeffs_out));
}
int p3r_fri_query(p3r_session* s, const uint32_t* indices, uint32_t n, uint32_t* out, size_t cap_words, size_t* n_words) {
    if (!s || !indices || !n_words) return P3R_ERR_INVALID_ARG;
    return fri_query_impl(s, indices, n, out, cap_words, n_words);
}
void p3r_session_free(p3r_session* s) {
    if (!s) return;
    ctx_wait(s->ctx);
    delete s;
}
int p3r_grind(p3r_ctx* ctx, const uint32_t state[16], const uint32_t* pending, uint32_t n_pending, uint32_t bits,
              uint32_t* witness_out) {
    if (!ctx || !state || !witness_out) return P3R_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    return DISPATCH(ctx, grind_impl<F>(ctx, state, pending, n_pending, bits, witness_out));
}
int p3r_prove(p3r_ctx* ctx, const p3r_prep* prep, const p3r_matrix_u32* traces, const uint32_t* const* public_values,
              uint32_t* proof_out, size_t cap_words, size_t* n_words) {
    if (!ctx || !prep || !traces || !n_words) return P3R_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    return DISPATCH(ctx, prove_impl<F>(ctx, prep, traces, public_values, proof_out, cap_words, n_words));
}
}  // extern "C" (template below)
template <class F>
static int traces_upload_impl(p3r_ctx* ctx, const p3r_prep* prep, const p3r_matrix_u32* traces,
                              const p3r_table_ops* tops, p3r_traces** out) {
    auto* t = new p3r_traces();
    t->ctx = ctx;
    ctx->arena.reset();
    size_t total = 0;
    for (const InstDev& d : prep->inst) total += ((((size_t)1 << d.log_h) * d.main_w * 4) + 255) & ~(size_t)255;
    if (cudaMalloc(&t->slab, std::max<size_t>(total, 256)) != cudaSuccess) {
        delete t;
        return P3R_ERR_OOM;
    }
    char* next = static_cast<char*>(t->slab);
    for (size_t i = 0; i < prep->inst.size(); i++) {
        const InstDev& d = prep->inst[i];
        const bool from_ops = tops && (tops[i].poseidon2 || tops[i].alu);
        if (!from_ops && (traces[i].height != (1u << d.log_h) || traces[i].width != d.main_w || !traces[i].data)) {
            set_err(ctx, "traces_upload: shape mismatch");
            p3r_traces_free(t);
            return P3R_ERR_INVALID_ARG;
        }
        size_t words = ((size_t)1 << d.log_h) * d.main_w;
        uint32_t* dm = reinterpret_cast<uint32_t*>(next);
        next += (words * 4 + 255) & ~(size_t)255;
        uint32_t* rm = from_ops ? reinterpret_cast<uint32_t*>(ctx->dstage) : arena_alloc<uint32_t>(ctx, words);
        if (!rm) {
            p3r_traces_free(t);
            return P3R_ERR_OOM;
        }
        t->d.push_back(dm);
        int rc = from_ops ? fill_from_ops<F>(ctx, d, tops[i], dm) : upload_matrix(ctx, traces[i], rm, dm);
        if (rc) {
            p3r_traces_free(t);
            return rc;
        }
    }
    ctx_wait(ctx);
    *out = t;
    return P3R_OK;
}
extern "C" {
static std::vector<p3r_table_ops> tops_from_p2(const p3r_prep* prep, const p3r_poseidon2_ops* const* p2_ops) {
    std::vector<p3r_table_ops> v(prep->inst.size(), p3r_table_ops{nullptr, nullptr});
    if (p2_ops)
        for (size_t i = 0; i < v.size(); i++) v[i].poseidon2 = p2_ops[i];
    return v;
}
int p3r_traces_upload_ops(p3r_ctx* ctx, const p3r_prep* prep, const p3r_matrix_u32* traces, const p3r_table_ops* table_ops,
                          p3r_traces** out) {
    if (!ctx || !prep || !traces || !out) return P3R_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    return DISPATCH(ctx, traces_upload_impl<F>(ctx, prep, traces, table_ops, out));
}
int p3r_traces_upload_ex(p3r_ctx* ctx, const p3r_prep* prep, const p3r_matrix_u32* traces, const p3r_poseidon2_ops* const* p2_ops,
                         p3r_traces** out) {
    if (!ctx || !prep || !traces || !out) return P3R_ERR_INVALID_ARG;
    std::vector<p3r_table_ops> tops = tops_from_p2(prep, p2_ops);
    return p3r_traces_upload_ops(ctx, prep, traces, tops.data(), out);
}
int p3r_traces_upload(p3r_ctx* ctx, const p3r_prep* prep, const p3r_matrix_u32* traces, p3r_traces** out) {
    return p3r_traces_upload_ex(ctx, prep, traces, nullptr, out);
}
int p3r_traces_download(p3r_ctx* ctx, const p3r_prep* prep, const p3r_traces* traces, uint32_t inst, uint32_t* out) {
    if (!ctx || !prep || !traces || !out || inst >= prep->inst.size()) return P3R_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    const InstDev& d = prep->inst[inst];
    uint32_t H = 1u << d.log_h;
    ctx->arena.reset();
    uint32_t* rm = arena_alloc<uint32_t>(ctx, (size_t)H * d.main_w);
    if (!rm) return P3R_ERR_OOM;
    dim3 grid((H + 31) / 32, (d.main_w + 31) / 32), block(32, 8);
    k_transpose_out<<<grid, block, 0, ctx->stream>>>(traces->d[inst], rm, H, d.main_w);
    LAUNCH_CHECK();
    CUDA_TRY(d2h_async(ctx, out, rm, (size_t)H * d.main_w * 4));
    CUDA_TRY(ctx_wait(ctx));
    return P3R_OK;
}
int p3r_prove_ex(p3r_ctx* ctx, const p3r_prep* prep, const p3r_matrix_u32* traces, const p3r_poseidon2_ops* const* p2_ops,
                 const uint32_t* const* public_values, uint32_t* proof_out, size_t cap_words, size_t* n_words) {
    if (!ctx || !prep || !traces || !n_words) return P3R_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    std::vector<p3r_table_ops> tops = tops_from_p2(prep, p2_ops);
    return DISPATCH(ctx, prove_impl<F>(ctx, prep, traces, public_values, proof_out, cap_words, n_words, nullptr, tops.data()));
}
int p3r_prove_ops(p3r_ctx* ctx, const p3r_prep* prep, const p3r_matrix_u32* traces, const p3r_table_ops* table_ops,
                  const uint32_t* const* public_values, uint32_t* proof_out, size_t cap_words, size_t* n_words) {
    if (!ctx || !prep || !traces || !n_words) return P3R_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    return DISPATCH(ctx, prove_impl<F>(ctx, prep, traces, public_values, proof_out, cap_words, n_words, nullptr, table_ops));
}
int p3r_traces_write_rows(p3r_ctx* ctx, const p3r_prep* prep, p3r_traces* traces, uint32_t inst, uint32_t row0, uint32_t n_rows,
                          const uint32_t* rows) {
    if (!ctx || !prep || !traces || !rows || inst >= prep->inst.size() || traces->d.size() != prep->inst.size())
        return P3R_ERR_INVALID_ARG;
    const InstDev& d = prep->inst[inst];
    const uint32_t H = 1u << d.log_h;
    if (n_rows == 0) return P3R_OK;
    if (row0 >= H || n_rows > H - row0 || (size_t)n_rows * d.main_w * 4 > ((size_t)1 << 20)) {
        set_err(ctx, "traces_write_rows: rows outside the table (or more than 1 MiB at once)");
        return P3R_ERR_INVALID_ARG;
    }
    cudaSetDevice(ctx->device);
    CUDA_TRY(ctx_wait(ctx));   // staging is about to be reused from its start
    ring_reset(ctx);
    const uint32_t* d_rows = (const uint32_t*)upload_small(ctx, rows, (size_t)n_rows * d.main_w * 4);
    if (!d_rows) return P3R_ERR_OOM;
    const uint32_t words = n_rows * d.main_w;
    k_scatter_rows<<<(words + 255) / 256, 256, 0, ctx->stream>>>(d_rows, traces->d[inst], H, d.main_w, row0, n_rows);
    LAUNCH_CHECK_C(KC_TRANSPOSE);
    CUDA_TRY(ctx_wait(ctx));   // the next session resets the staging ring
    return P3R_OK;
}
void p3r_traces_free(p3r_traces* t) {
    if (!t) return;
    ctx_wait(t->ctx);
    cudaFree(t->slab);
    delete t;
}
int p3r_prove_resident(p3r_ctx* ctx, const p3r_prep* prep, const p3r_traces* traces, const uint32_t* const* public_values,
                       uint32_t* proof_out, size_t cap_words, size_t* n_words) {
    if (!ctx || !prep || !traces || !n_words || traces->d.size() != prep->inst.size()) return P3R_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    return DISPATCH(ctx, prove_impl<F>(ctx, prep, nullptr, public_values, proof_out, cap_words, n_words, traces));
}
int p3r_set_specialization(p3r_ctx* ctx, int enable) {
    if (!ctx) return P3R_ERR_INVALID_ARG;
    // bit 0: build-time specialised quotient kernels; bit 1 set = DISABLE the whole-column LDE kernels (tile kernel only),
    // so enable = 1 / 0 keep their old meaning and the parity tests can cross-check both LDE paths.
    ctx->use_spec = (enable & 1) != 0;
    ctx->use_col_ntt = (enable & 2) == 0;
    ctx->dev_fri_transcript = (enable & 4) == 0;
    ctx->use_hash_queue = (enable & 8) != 0;
    ctx->use_grouped_interp = (enable & 16) == 0;
    return P3R_OK;
}
int p3r_timer_start(p3r_ctx* ctx) {
    if (!ctx) return P3R_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    if (!ctx->timer_ev[0]) {
        cudaEventCreate(&ctx->timer_ev[0]);
        cudaEventCreate(&ctx->timer_ev[1]);
    }
    CUDA_TRY(ctx_wait(ctx));
    CUDA_TRY(cudaEventRecord(ctx->timer_ev[0], ctx->stream));
    return P3R_OK;
}
int p3r_timer_stop(p3r_ctx* ctx, float* ms_out) {
    if (!ctx || !ms_out || !ctx->timer_ev[0]) return P3R_ERR_INVALID_ARG;
    CUDA_TRY(cudaEventRecord(ctx->timer_ev[1], ctx->stream));
    CUDA_TRY(cudaEventSynchronize(ctx->timer_ev[1]));
    CUDA_TRY(cudaEventElapsedTime(ms_out, ctx->timer_ev[0], ctx->timer_ev[1]));
    return P3R_OK;
}
void* p3r_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void p3r_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
int p3r_set_kernel_timing(p3r_ctx* ctx, uint32_t class_mask) {
    if (!ctx) return P3R_ERR_INVALID_ARG;
    kstats_collect(ctx);
    ctx->time_mask = class_mask;
    return P3R_OK;
}
int p3r_reset_kernel_stats(p3r_ctx* ctx) {
    if (!ctx) return P3R_ERR_INVALID_ARG;
    kstats_collect(ctx);
    ctx->kstats = KernelStats();
    return P3R_OK;
}
int p3r_kernel_stats(p3r_ctx* ctx, const char** names_out, double* ms_out, uint64_t* launches_out, uint64_t* bytes_out,
                     uint32_t cap, uint32_t* n_out) {
    if (!ctx || !n_out) return P3R_ERR_INVALID_ARG;
    kstats_collect(ctx);
    for (uint32_t k = 0; k < KC_COUNT && k < cap; k++) {
        if (names_out) names_out[k] = KCLASS_NAMES[k];
        if (ms_out) ms_out[k] = ctx->kstats.ms[k];
        if (launches_out) launches_out[k] = ctx->kstats.launches[k];
        if (bytes_out) bytes_out[k] = ctx->kstats.bytes[k];
    }
    *n_out = KC_COUNT;
    return P3R_OK;
}
int p3r_kernel_perms(p3r_ctx* ctx, uint64_t* perms_out, uint32_t cap) {
    if (!ctx || !perms_out) return P3R_ERR_INVALID_ARG;
    for (uint32_t k = 0; k < KC_COUNT && k < cap; k++) perms_out[k] = ctx->kstats.perms[k];
    return P3R_OK;
}
int p3r_coset_lde(p3r_ctx* ctx, const p3r_matrix_u32* in, uint32_t log_blowup, uint32_t* out) {
    if (!ctx || !in || !out) return P3R_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    return DISPATCH(ctx, coset_lde_host_impl<F>(ctx, in, log_blowup, out));
}
int p3r_mmcs_commit(p3r_ctx* ctx, uint32_t n_mats, const p3r_matrix_u32* mats, uint32_t* cap_out) {
    if (!ctx || !mats || !cap_out || !n_mats) return P3R_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    return DISPATCH(ctx, mmcs_commit_host_impl<F>(ctx, n_mats, mats, cap_out));
}
int p3r_poseidon2_permute(p3r_ctx* ctx, uint32_t* states, uint32_t n) {
    if (!ctx || !states) return P3R_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    return DISPATCH(ctx, permute_host_impl<F>(ctx, states, n));
}
}  // extern "C" (template below)
// Device-resident benchmark of one FRI commit round on a synthetic EF vector of 2^log_len elements: fold by 2^log_arity,
// then Merkle-commit the folded vector's arity-wide rows.
template <class F>
static int bench_fri_round_impl(p3r_ctx* ctx, uint32_t log_len, uint32_t log_arity, uint32_t iters, uint64_t seed, float* ms_out) {
    ctx->arena.reset();
    ring_reset(ctx);
    if (log_arity < 1 || log_arity > 4 || log_len < 2 * log_arity + ctx->fri.cap_height) return P3R_ERR_INVALID_ARG;
    TRY(ensure_twiddles<F>(ctx, log_len));
    const size_t L = (size_t)1 << log_len;
    const uint32_t out_log = log_len - log_arity, rows_log = out_log - log_arity;
    Ext4* in = arena_alloc<Ext4>(ctx, L);
    Ext4* out = arena_alloc<Ext4>(ctx, L >> log_arity);
    uint32_t* dg = arena_alloc<uint32_t>(ctx, tree_digest_words(rows_log));
    if (!in || !out || !dg) return P3R_ERR_OOM;
    k_fill_random<F><<<(unsigned)((L * 4 + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<uint32_t*>(in), L * 4, seed);
    LAUNCH_CHECK();
    Ext4 beta{{to_monty<F>(3), to_monty<F>(5), to_monty<F>(7), to_monty<F>(11)}};
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventCreate(&e2);
    float t_fold = 0, t_tree = 0;
    for (uint32_t it = 0; it <= iters; it++) {   // iteration 0 = warm-up
        size_t pin_mark = ctx->pin_used;
        cudaEventRecord(e0, ctx->stream);
        k_fri_fold<F><<<(unsigned)(((L >> log_arity) + 127) / 128), 128, 0, ctx->stream>>>(in, out, log_len, log_arity, beta, nullptr,
                                                                                          nullptr, ctx->inv2_m, ctx->tw, ctx->logT, ctx->w_m);
        LAUNCH_CHECK();
        cudaEventRecord(e1, ctx->stream);
        TRY(build_tree<F>(ctx, reinterpret_cast<const uint32_t*>(out), 4u << log_arity, rows_log, dg,
                          [](uint32_t) { return (const uint32_t*)nullptr; }));
        cudaEventRecord(e2, ctx->stream);
        CUDA_TRY(cudaEventSynchronize(e2));
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, e0, e1);
        cudaEventElapsedTime(&b, e1, e2);
        if (it) t_fold += a, t_tree += b;
        ctx->pin_used = pin_mark;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    ms_out[0] = t_fold / iters;
    ms_out[1] = t_tree / iters;
    return P3R_OK;
}
extern "C" {
int p3r_bench_fri_round(p3r_ctx* ctx, uint32_t log_len, uint32_t log_arity, uint32_t iters, uint64_t seed, float* times_ms_out) {
    if (!ctx || !times_ms_out || !iters) return P3R_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    return DISPATCH(ctx, bench_fri_round_impl<F>(ctx, log_len, log_arity, iters, seed, times_ms_out));
}
int p3r_poseidon2_run_chains(p3r_ctx* ctx, const p3r_poseidon2_chain_ops* ops, uint32_t* inputs_out, uint32_t* outputs_out) {
    if (!ctx || !ops || !inputs_out || !outputs_out) return P3R_ERR_INVALID_ARG;
    if (ops->n_rows && (!ops->new_start || !ops->merkle_path || !ops->mmcs_bit || !ops->witness_mask || !ops->values)) return P3R_ERR_INVALID_ARG;
    if ((size_t)ops->n_rows * 64 > ctx->pin_out_size / 2) {
        set_err(ctx, "run_chains: at most " + std::to_string(ctx->pin_out_size / 128) + " rows per call");
        return P3R_ERR_INVALID_ARG;
    }
    cudaSetDevice(ctx->device);
    return DISPATCH(ctx, run_chains_impl<F>(ctx, ops, inputs_out, outputs_out));
}
int p3r_bench_commit_multi(p3r_ctx* ctx, uint32_t n_mats, const uint32_t* log_heights, const uint32_t* widths, uint32_t iters,
                           uint64_t seed, float* times_ms_out) {
    if (!ctx || !times_ms_out || !iters || !n_mats || !log_heights || !widths) return P3R_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    return DISPATCH(ctx, bench_commit_multi_impl<F>(ctx, n_mats, log_heights, widths, iters, seed, times_ms_out));
}
int p3r_bench_commit(p3r_ctx* ctx, uint32_t log_height, uint32_t width, uint32_t iters, uint64_t seed, float* times_ms_out) {
    if (!ctx || !times_ms_out || !iters) return P3R_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    return DISPATCH(ctx, bench_commit_impl<F>(ctx, log_height, width, iters, seed, times_ms_out));
}
int p3r_last_phase_times(p3r_ctx* ctx, const char** names_out, float* ms_out, uint32_t cap, uint32_t* n_out) {
    if (!ctx || !n_out) return P3R_ERR_INVALID_ARG;
    ctx->phase_names_blob.clear();
    uint32_t n = 0;
    for (auto& kv : ctx->phase_times) {
        if (n < cap && ms_out) ms_out[n] = kv.second;
        ctx->phase_names_blob += kv.first;
        ctx->phase_names_blob.push_back('\0');
        n++;
    }
    ctx->phase_names_blob.push_back('\0');
    if (names_out) *names_out = ctx->phase_names_blob.c_str();
    *n_out = n;
    return P3R_OK;
}
uint64_t p3r_launch_count(const p3r_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"

// ---- host-only Poseidon2 (no CUDA call): the transcript's permutation behind a handle, for host code that needs the same
// hash the prover uses — the circuit runner's Poseidon2 rows (SURVEY.md §8f item 4 when no GPU batch is worth it) and the
// synthetic-workload generator. States are Montgomery words, like p3r_poseidon2_permute.
struct p3r_host_hasher {
    int field_id = 0;
    Poseidon2Consts k{};
    HostPermuteFn fn = nullptr;
};
extern "C" {
int p3r_host_hasher_create(const p3r_field_desc* field, const p3r_poseidon2_consts* p2, p3r_host_hasher** out) {
    if (!field || !p2 || !out || !p2->external_rc || !p2->internal_rc || !p2->internal_diag) return P3R_ERR_INVALID_ARG;
    uint32_t want_p = field->field_id == P3R_FIELD_KOALABEAR ? KoalaBear::P : BabyBear::P;
    if (field->field_id > 1 || field->p != want_p) return P3R_ERR_UNSUPPORTED;
    uint32_t rp = field->field_id == P3R_FIELD_KOALABEAR ? KoalaBear::ROUNDS_P : BabyBear::ROUNDS_P;
    uint32_t sb = field->field_id == P3R_FIELD_KOALABEAR ? KoalaBear::SBOX : BabyBear::SBOX;
    if (p2->width != 16 || p2->rounds_f != 8 || p2->rounds_p != rp || p2->sbox_degree != sb) return P3R_ERR_UNSUPPORTED;
    auto* h = new p3r_host_hasher();
    h->field_id = (int)field->field_id;
    std::memcpy(h->k.ext_rc, p2->external_rc, 8 * 16 * 4);
    std::memset(h->k.int_rc, 0, sizeof h->k.int_rc);
    std::memcpy(h->k.int_rc, p2->internal_rc, rp * 4);
    std::memcpy(h->k.diag, p2->internal_diag, 16 * 4);
    h->k.zero = 0;
    bool fast = true;
    for (int i = 0; i < 16; i++) {
        uint32_t want = h->field_id == 0 ? to_monty<KoalaBear>(diag_spec_canonical<KoalaBear>(i))
                                         : to_monty<BabyBear>(diag_spec_canonical<BabyBear>(i));
        fast = fast && h->k.diag[i] == want;
    }
    h->k.fast_diag = fast ? 1u : 0u;
    h->fn = h->field_id == 0 ? pick_host_permute<KoalaBear>(h->k) : pick_host_permute<BabyBear>(h->k);
    *out = h;
    return P3R_OK;
}
int p3r_host_hasher_permute(const p3r_host_hasher* h, uint32_t* states, size_t n) {
    if (!h || (!states && n)) return P3R_ERR_INVALID_ARG;
    for (size_t i = 0; i < n; i++) h->fn(states + 16 * i, h->k);
    return P3R_OK;
}
void p3r_host_hasher_free(p3r_host_hasher* h) { delete h; }
}  // extern "C"
