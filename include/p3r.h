/*
 * p3r.h — C ABI of the B200-native batch-STARK prover behind Plonky3-recursion's
 *         `BatchStarkProver::prove_all_tables` / `ProverData::from_airs_and_degrees`.
 *
 * This is the drop-in boundary. A Rust `GpuBatchStarkProver<SC>` keeps the
 * reference signatures
 *     prove_all_tables(&self, &Traces<EF>, &CircuitProverData<SC>)      circuit-prover/src/batch_stark_prover.rs:1203-1222
 *     ProverData::from_airs_and_degrees(..)                              recursion/src/recursion.rs:376,487,737
 * and calls the functions below through a `-sys` crate (see INTEGRATION.md for the binding stub).
 * Fiat–Shamir stays on the host (`SC::Challenger`); every function that needs a challenge takes it
 * as an argument and every function that produces transcript material returns it in caller memory.
 *
 * Conventions
 *  - return value 0 = OK, non-zero = p3r_status (mapped to BatchStarkProverError on the Rust side,
 *    circuit-prover/src/batch_stark_prover.rs:786-808); p3r_last_error() gives a message.
 *  - every field element is a u32 in Montgomery form (R = 2^32), little-endian, the in-memory form of
 *    p3's MontyField31; an extension element is 4 consecutive coefficients (BinomialExtensionField<F,4>).
 *  - host matrices are row-major (p3 RowMajorMatrix); the library transposes to its column-major
 *    device layout on upload. Caller owns all host buffers; handles own device memory.
 *  - one session = one CUDA stream; a ctx is bound to one device; sessions of one ctx are not
 *    concurrently callable.
 *  - there is NO CPU fallback: every entry point fails with P3R_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef P3R_H
#define P3R_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct p3r_ctx p3r_ctx;
typedef struct p3r_prep p3r_prep;
typedef struct p3r_session p3r_session;
typedef struct p3r_traces p3r_traces;

typedef enum {
    P3R_OK = 0,
    P3R_ERR_INVALID_ARG = 1,   /* shape / parameter error            -> BatchStarkProverError::InvalidTableShape-like */
    P3R_ERR_CUDA = 2,          /* CUDA runtime error / no device     -> BatchStarkProverError::Backend               */
    P3R_ERR_OOM = 3,           /* device or host allocation failed                                                    */
    P3R_ERR_STATE = 4,         /* phase called out of order                                                           */
    P3R_ERR_UNSUPPORTED = 5,   /* field / width / option not built                                                    */
    P3R_ERR_BUFFER = 6,        /* caller buffer too small (needed size is written to the size out-param)              */
    P3R_ERR_POW = 7            /* no proof-of-work witness found                                                      */
} p3r_status;

enum { P3R_FIELD_KOALABEAR = 0, P3R_FIELD_BABYBEAR = 1 };

/* Field description. `w` is the binomial constant x^4 = w that the reference obtains from
 * ExtractBinomialW::extract_w (circuit-prover/src/field_params.rs:34-41) and records in the proof
 * (batch_stark_prover.rs:627). Values are CANONICAL (not Montgomery) integers. */
typedef struct {
    uint32_t field_id;      /* P3R_FIELD_*                                    */
    uint32_t p;             /* modulus, must match field_id                   */
    uint32_t w;             /* binomial non-residue (3 KoalaBear, 11 BabyBear) */
    uint32_t generator;     /* F::GENERATOR = coset shift of every LDE        */
} p3r_field_desc;

/* Poseidon2 parameters, width 16 (context permutation) or 24 (leaf hasher); round constants are injected, Montgomery form.
 * Reference: poseidon2-circuit-air/src/public_types.rs:48-53,99-104,220-226,272-278. */
typedef struct {
    uint32_t width;                 /* 16 (24 for p3r_ctx_set_leaf_hasher / p3r_poseidon2_permute_w) */
    uint32_t sbox_degree;           /* 3 (KoalaBear) or 7 (BabyBear)                */
    uint32_t rounds_f;              /* 8 (4 initial + 4 terminal)                   */
    uint32_t rounds_p;              /* 20 (KoalaBear) or 13 (BabyBear)              */
    const uint32_t* external_rc;    /* rounds_f * width words: initial rounds then terminal rounds */
    const uint32_t* internal_rc;    /* rounds_p words                                */
    const uint32_t* internal_diag;  /* width words V: s_i <- V_i * s_i + sum(s)      */
} p3r_poseidon2_consts;

/* FRI / MMCS parameters (recursion/examples/common/mod.rs:464-486, circuit-prover/src/config.rs:129-136). */
typedef struct {
    uint32_t log_blowup;
    uint32_t log_final_poly_len;
    uint32_t max_log_arity;
    uint32_t num_queries;
    uint32_t commit_pow_bits;
    uint32_t query_pow_bits;
    uint32_t cap_height;            /* Merkle cap has 2^cap_height digests */
} p3r_fri_params;

/* Row-major host matrix of Montgomery words (p3 RowMajorMatrix<Val>). */
typedef struct {
    const uint32_t* data;
    uint32_t height;
    uint32_t width;
} p3r_matrix_u32;

/* ------------------------------------------------------------------------------------------------
 * Constraint bytecode. The Rust side compiles the (base, ext) SymbolicExpression DAGs returned by
 * p3_batch_stark::symbolic::get_symbolic_constraints (the call the recursive verifier makes,
 * recursion/src/traits/air.rs:160) into this SSA form once per circuit shape; node kinds follow
 * circuit/src/symbolic/compiler.rs:86-118,183-189. Two register files: base slots (1 word) and
 * extension slots (4 words). `a`/`b` name slots unless stated otherwise.
 * Folding follows recursion/src/traits/air.rs:170-181: acc = acc*alpha + c over all base constraints
 * in order, then all extension constraints; ASSERT_*'s `dst` is the position in that combined order.
 * ---------------------------------------------------------------------------------------------- */
typedef enum {
    P3R_OP_B_MAIN = 0,    /* B[dst] = main[col=a][row + b]          b in {0,1}                  */
    P3R_OP_B_PREP = 1,    /* B[dst] = preprocessed[col=a][row + b]                               */
    P3R_OP_B_PUB = 2,     /* B[dst] = public_values[a]                                           */
    P3R_OP_B_SEL = 3,     /* B[dst] = selector a: 0 is_first_row, 1 is_last_row, 2 is_transition */
    P3R_OP_B_CONST = 4,   /* B[dst] = a (Montgomery immediate)                                   */
    P3R_OP_B_ADD = 5,
    P3R_OP_B_SUB = 6,
    P3R_OP_B_MUL = 7,
    P3R_OP_B_NEG = 8,
    P3R_OP_E_PERM = 16,   /* E[dst] = permutation EF column a at row + b                         */
    P3R_OP_E_CHAL = 17,   /* E[dst] = challenge a: 2c = bus prefix of lookup c, 2c+1 = beta      */
    P3R_OP_E_PVAL = 18,   /* E[dst] = permutation value a (the AIR's LogUp terminal)             */
    P3R_OP_E_CONST = 19,  /* E[dst] = ext_consts[a]                                              */
    P3R_OP_E_FROMB = 20,  /* E[dst] = lift(B[a])                                                 */
    P3R_OP_E_ADD = 21,
    P3R_OP_E_SUB = 22,
    P3R_OP_E_MUL = 23,
    P3R_OP_E_NEG = 24,
    P3R_OP_E_MULB = 25,   /* E[dst] = E[a] * B[b]                                                */
    P3R_OP_E_ADDB = 26,   /* E[dst] = E[a] + B[b]                                                */
    P3R_OP_E_SUBB = 27,   /* E[dst] = E[a] - B[b]                                                */
    P3R_OP_ASSERT_B = 32, /* constraint #dst (fold order) = B[a]                                 */
    P3R_OP_ASSERT_E = 33, /* constraint #dst (fold order) = E[a]                                 */
    P3R_OP_OUT_B = 34     /* lookup-input programs: out[dst] = B[a]                              */
} p3r_opcode;

typedef struct {
    uint32_t op, dst, a, b;
} p3r_insn;

typedef struct {
    const p3r_insn* insns;
    uint32_t n_insns;
    uint32_t n_base_slots;
    uint32_t n_ext_slots;
    const uint32_t* ext_consts;     /* 4 Montgomery words each */
    uint32_t n_ext_consts;
    uint32_t n_constraints;         /* base + ext, constraint programs */
    uint32_t n_outputs;             /* lookup-input programs           */
} p3r_program;

/* LogUp (`WitnessChecks` bus) description of one AIR after pack_same_bus
 * (circuit-prover/src/batch_stark_prover.rs:925-941). One p3r_lookup = one fraction column
 * (permutation column c+1, recursion/src/verifier/batch_stark.rs:902-912); its interactions are
 * (multiplicity, tuple) pairs whose values the `lookup_inputs` program writes with OUT_B. */
typedef struct {
    uint32_t mult_out;          /* OUT_B index of the multiplicity             */
    uint32_t elem_out_first;    /* OUT_B index of tuple element 0               */
    uint32_t n_elems;           /* tuple length; elements are consecutive OUT_B */
} p3r_interaction;

typedef struct {
    uint32_t bus;               /* global bus id (recursion/src/verifier/batch_stark.rs:1055-1083) */
    uint32_t first_interaction;
    uint32_t n_interactions;
} p3r_lookup;

/* One table (AIR instance). Order of instances is the reference's
 * [Const, Public, ALU, NPOs...] (circuit-prover/src/batch_stark_prover.rs:1493-1519). */
typedef struct {
    uint32_t log_height;            /* log2 of the padded trace height                                  */
    uint32_t main_width;
    uint32_t prep_width;            /* 0 = no preprocessed columns                                      */
    uint32_t n_public;
    uint32_t log_quotient_chunks;
    uint32_t uses_next_row;         /* main trace is opened at zeta*g (recursion/src/traits/air.rs:103-109) */
    p3r_program constraints;        /* AIR + LogUp constraints                                          */
    p3r_program lookup_inputs;      /* n_outputs = values consumed by `interactions`                    */
    const p3r_lookup* lookups;
    uint32_t n_lookups;
    const p3r_interaction* interactions;
    uint32_t n_interactions;
} p3r_instance_desc;

/* Operation list of the Poseidon2 table (the fields of `Poseidon2CircuitRow`, circuit/src/ops/poseidon2_perm/trace.rs:94-124,
 * that determine the MAIN trace; new_start / merkle_path are read from the committed preprocessed columns). When given for an
 * instance instead of a matrix, the library fills the full round-state trace on the device
 * (replaces Poseidon2CircuitAir::generate_trace_rows, poseidon2-circuit-air/src/air.rs:280-520). */
typedef struct {
    uint32_t n_ops;
    const uint32_t* input_values;      /* n_ops * 16 Montgomery words                                  */
    const uint8_t* mmcs_bit;           /* n_ops                                                        */
    const uint32_t* mmcs_index_sum;    /* n_ops Montgomery words (value where the accumulator restarts) */
} p3r_poseidon2_ops;

/* Operation list of the ALU table: the schedule (static per circuit shape: AluAir::compute_schedule,
 * circuit-prover/src/air/alu_air.rs:349-463) and the operand values the runner produced. The library scatters them into
 * the main trace on the device (replaces AluAir::trace_to_matrix, alu_air.rs:497-608), including the packed-Horner
 * intermediates, (a_t, c_t) operands and b^2 of lane 0. Only D = 4 with horner k_max = 4 (the recursion-layer packing). */
typedef struct {
    uint32_t lanes, d, k_max;
    uint32_t n_slots;              /* schedule slots in use (slot = row * lanes + lane); the rest of the table is padding */
    const uint32_t* slot_kind;     /* per slot: 0 = separator / empty, 1 = one operation, k >= 2 = packed Horner of k operations */
    const uint32_t* slot_first;    /* per slot: index of its (first) operation */
    uint32_t n_ops;
    const uint32_t* values;        /* n_ops * 4 * d Montgomery words: a, b, c, out */
} p3r_alu_ops;

/* Where the main trace of one instance comes from: the matrix passed alongside (both pointers NULL) or an operation list. */
typedef struct {
    const p3r_poseidon2_ops* poseidon2;
    const p3r_alu_ops* alu;
} p3r_table_ops;

/* ---------------------------------------------------------------------------------------------- */

/* Library/ABI version and a description of the build (arch, fields). */
uint32_t p3r_abi_version(void);
const char* p3r_build_info(void);

/* Create a prover context on CUDA device `device`. Replaces building `StarkConfig`/`MyPcs`
 * (recursion/examples/common/mod.rs:464-486; circuit-prover/src/config.rs:92-137). */
int p3r_ctx_create(int device, const p3r_field_desc* field, const p3r_poseidon2_consts* poseidon2,
                   const p3r_fri_params* fri, p3r_ctx** out);
void p3r_ctx_destroy(p3r_ctx* ctx);
const char* p3r_last_error(const p3r_ctx* ctx);

/* Protocol conventions that live in the crates.io p3-* 0.6 crates and NOT in the reference tree ([P3-EXT]): each is
 * a named, run-time selectable choice in the library and, identically, in the oracle, so that pinning against real reference
 * output (tools/ref_golden) is a flag flip, not a rewrite. Defaults (all zero) are the choices DESIGN.md §2 argues for.
 * LogUp denominator of a tuple (f_0 .. f_{n-1}) on a bus:  bus_prefix + s * sum_k beta^{e(k)} * f_k  with
 *   s = -1 if logup_negate else +1;   e(k) = first_power + (logup_descending ? n - 1 - k : k).
 * The constraint bytecode the caller supplies must encode the same choice (symbolic.LOGUP_CONVENTIONS on the Python side). */
typedef struct {
    uint32_t logup_negate;        /* 0: prefix + sum, 1: prefix - sum                                  */
    uint32_t logup_first_power;   /* 0: powers start at beta^0, 1: at beta^1                           */
    uint32_t logup_descending;    /* 0: f_0 gets the lowest power, 1: the highest (all tuples of one length) */
} p3r_conventions;
int p3r_ctx_set_conventions(p3r_ctx* ctx, const p3r_conventions* conv);

/* Uni-STARK mode (the base layer of recursive_keccak, `p3_uni_stark::prove(&config_0, &keccak_air, trace, &pis)`,
 * recursion/examples/recursive_keccak.rs:520-530): the one-shot provers (p3r_prove*) prove ONE table without lookups with the
 * uni-stark transcript head — degree_bits, degree_bits - is_zk, preprocessed width as single base elements, trace commitment,
 * preprocessed commitment, public values (restated in-tree at recursion/src/types/challenges.rs:44-54,100-140) — instead of the
 * batch head; everything after it (alpha, quotient, zeta, openings at zeta and zeta*g, FRI) is shared with the batch prover, which
 * is why the same kernels serve the wide single-table shape (KeccakAir: ~2 600 columns). The proof blob layout is unchanged. */
int p3r_ctx_set_uni_stark(p3r_ctx* ctx, int on);

/* Width-24 leaf hashing: with `w24` (width 24 constants of the context's field: 21 / 23 partial rounds for BabyBear / KoalaBear,
 * circuit/src/ops/poseidon2_perm/config.rs:77-86,124-133) every MMCS leaf row — trace, quotient and FRI commit-phase matrices — is
 * hashed with PaddingFreeSponge<Perm24, 24, 16, 8> instead of the width-16 sponge; the 2-to-1 compression and the challenger stay on
 * the context's width-16 permutation (the `Config<F, PermHash, PermCompress, HASH_PERM_WIDTH, ..>` shape of
 * circuit-prover/src/config.rs:59-74 with PermHash = Poseidon2<24>). NULL switches back to width 16. Call between proofs. */
int p3r_ctx_set_leaf_hasher(p3r_ctx* ctx, const p3r_poseidon2_consts* w24);
/* Poseidon2 permutation of n states of consts->width (16 or 24) words, in place, with the GIVEN constants (isolated kernel). */
int p3r_poseidon2_permute_w(p3r_ctx* ctx, const p3r_poseidon2_consts* consts, uint32_t* states, uint32_t n);

/* Replaces ProverData::from_airs_and_degrees (recursion/src/recursion.rs:376): uploads the
 * instance descriptions (bytecode, lookups), LDEs + commits all preprocessed matrices into one
 * global MMCS tree that stays device-resident, and writes its cap (8 << cap_height words).
 * `prep[i].data` may be NULL when descs[i].prep_width == 0. `cap_out` is untouched when no
 * instance has preprocessed columns (then *has_prep_out = 0). */
int p3r_prep_commit(p3r_ctx* ctx, uint32_t n_inst, const p3r_instance_desc* descs,
                    const p3r_matrix_u32* prep, p3r_prep** out, uint32_t* cap_out, uint32_t* has_prep_out);
void p3r_prep_free(p3r_prep* prep);

/* Phase-stepped proving session; replaces p3_batch_stark::prove_batch
 * (call site circuit-prover/src/batch_stark_prover.rs:1595). Phases must be called in the order
 * they are declared. `public_values[i]` has descs[i].n_public words (may be NULL when 0). */
int p3r_prove_begin(p3r_ctx* ctx, const p3r_prep* prep, const p3r_matrix_u32* traces,
                    const uint32_t* const* public_values, p3r_session** out);
/* LDE + MMCS commit of all main traces. cap_out: 8 << cap_height words. */
int p3r_commit_main(p3r_session* s, uint32_t* cap_out);
/* LogUp permutation traces for (alpha, beta), LDE + commit. terminals_out: 4 words per instance
 * that has lookups, in instance order. No-op returning P3R_OK when no instance has lookups. */
int p3r_commit_perm(p3r_session* s, const uint32_t alpha[4], const uint32_t beta[4], uint32_t* cap_out,
                    uint32_t* terminals_out);
/* Quotient evaluation with folding challenge alpha, chunk split, LDE + commit. */
int p3r_commit_quotient(p3r_session* s, const uint32_t alpha[4], uint32_t* cap_out);
/* Out-of-domain openings at zeta (and zeta*g_i). Layout of opened_out, per instance i in order:
 *   main_local[main_width*4], main_next[main_width*4] (iff uses_next_row),
 *   prep_local[prep_width*4], prep_next[prep_width*4] (iff prep_width>0),
 *   perm_local[aux*4*4], perm_next[aux*4*4] (iff n_lookups>0; aux = n_lookups+1; one EF per flattened base column),
 *   quotient_chunks[2^log_qc][4*4]
 * (every opened value is an EF element = 4 words). *n_words receives the size. */
int p3r_open(p3r_session* s, const uint32_t zeta[4], uint32_t* opened_out, size_t cap_words, size_t* n_words);
/* Reduced openings per height with alpha_fri; fixes the arity schedule. log_arities_out: up to 32 entries. */
int p3r_fri_begin(p3r_session* s, const uint32_t alpha_fri[4], uint32_t* n_rounds_out, uint32_t* log_arities_out);
/* Commit FRI round `round` (matrix of folded-height rows x arity EF), write its cap. */
int p3r_fri_commit(p3r_session* s, uint32_t round, uint32_t* cap_out);
/* Fold round `round` with beta and roll in the reduced opening of the folded height. */
int p3r_fri_fold(p3r_session* s, uint32_t round, const uint32_t beta[4]);
/* Final polynomial coefficients: 4 << log_final_poly_len words. */
int p3r_fri_final_poly(p3r_session* s, uint32_t* coeffs_out);
/* Query openings for `n` indices (each < 2^log_global_max_height). Blob layout, per query:
 *   for each input round in [main, quotient, preprocessed?, permutation?]:
 *       for each matrix of the round in commit order: the opened row (width words)
 *       Merkle path: (log_max_height_of_round - cap_height) digests of 8 words, leaf to cap
 *   for each FRI round r: (2^log_arity_r - 1) sibling EF values (4 words each, group order with own slot removed),
 *       Merkle path of (log_folded_height_r - cap_height) digests */
int p3r_fri_query(p3r_session* s, const uint32_t* indices, uint32_t n, uint32_t* out, size_t cap_words,
                  size_t* n_words);
void p3r_session_free(p3r_session* s);

/* DuplexChallenger::grind on the device (p3-challenger GrindingChallenger::grind; check restated at
 * recursion/src/challenger/circuit.rs:409-430). `state` = 16 sponge words, `pending` = the n_pending (<8)
 * words in the challenger's input buffer. Returns the SMALLEST canonical witness w such that after
 * observe(w) the next sample has `bits` low zero bits (deterministic, unlike a parallel find_any). */
int p3r_grind(p3r_ctx* ctx, const uint32_t state[16], const uint32_t* pending, uint32_t n_pending, uint32_t bits,
              uint32_t* witness_out);

/* One-shot prove: runs the host DuplexChallenger transcript (host C++ mirror of BatchStarkProver::prove,
 * circuit-prover/src/batch_stark_prover.rs:1275-1642) over the phases above and writes the flat proof
 * blob described in DESIGN.md §"Proof blob". */
int p3r_prove(p3r_ctx* ctx, const p3r_prep* prep, const p3r_matrix_u32* traces,
              const uint32_t* const* public_values, uint32_t* proof_out, size_t cap_words, size_t* n_words);

/* Variants taking, per instance, either a matrix (p2_ops[i] == NULL) or a Poseidon2 operation list (p2_ops[i] != NULL, the
 * matrix entry is ignored): the Poseidon2 table is then generated on the device. `p2_ops` itself may be NULL. */
int p3r_prove_ex(p3r_ctx* ctx, const p3r_prep* prep, const p3r_matrix_u32* traces, const p3r_poseidon2_ops* const* p2_ops,
                 const uint32_t* const* public_values, uint32_t* proof_out, size_t cap_words, size_t* n_words);
int p3r_traces_upload_ex(p3r_ctx* ctx, const p3r_prep* prep, const p3r_matrix_u32* traces,
                         const p3r_poseidon2_ops* const* p2_ops, p3r_traces** out);
/* General form: table_ops (n_instances entries, or NULL) says per instance whether the trace is the matrix or an operation
 * list expanded on the device (Poseidon2 and ALU tables). */
int p3r_prove_ops(p3r_ctx* ctx, const p3r_prep* prep, const p3r_matrix_u32* traces, const p3r_table_ops* table_ops,
                  const uint32_t* const* public_values, uint32_t* proof_out, size_t cap_words, size_t* n_words);
int p3r_traces_upload_ops(p3r_ctx* ctx, const p3r_prep* prep, const p3r_matrix_u32* traces, const p3r_table_ops* table_ops,
                          p3r_traces** out);
/* Debug/parity helper: download the (row-major) main trace of instance `inst` from device-resident traces. */
int p3r_traces_download(p3r_ctx* ctx, const p3r_prep* prep, const p3r_traces* traces, uint32_t inst, uint32_t* out);

/* Device-resident traces: upload (and transpose to the column-major device layout) once, prove many times. This is the
 * path a caller uses when the trace builders already ran on the device or when the same traces are proved repeatedly;
 * bench.py uses it for the device-resident `value` next to the host-buffer `e2e` number. */
int p3r_traces_upload(p3r_ctx* ctx, const p3r_prep* prep, const p3r_matrix_u32* traces, p3r_traces** out);
void p3r_traces_free(p3r_traces* t);
/* Overwrite rows [row0, row0 + n_rows) of the device-resident main trace of instance `inst` with `rows` (row-major,
 * n_rows x main_width Montgomery words, host memory). This is how the per-proof part of an otherwise cached witness reaches the
 * device: in the aggregation tree (recursion/examples/recursive_aggregation.rs:676-704) a node's Public table carries values
 * taken from its two child proofs while the rest of its traces keep their shape. Synchronous (returns after the copy). */
int p3r_traces_write_rows(p3r_ctx* ctx, const p3r_prep* prep, p3r_traces* traces, uint32_t inst, uint32_t row0,
                          uint32_t n_rows, const uint32_t* rows);
int p3r_prove_resident(p3r_ctx* ctx, const p3r_prep* prep, const p3r_traces* traces, const uint32_t* const* public_values,
                       uint32_t* proof_out, size_t cap_words, size_t* n_words);
/* Pinned host memory for trace/proof buffers (cudaHostAlloc); plain malloc'd buffers also work, just slower to copy. */
void* p3r_host_alloc(size_t bytes);
void p3r_host_free(void* p);

/* ---- isolated kernels (LDE + Merkle + FRI-fold sweep). Host pointers. ---- */
/* Coset LDE of a row-major height x width matrix -> row-major (height<<log_blowup) x width, rows bit-reversed
 * (TwoAdicFriPcs::commit -> Radix2DitParallel::coset_lde_batch, circuit-prover/src/config.rs:17,131). */
int p3r_coset_lde(p3r_ctx* ctx, const p3r_matrix_u32* in, uint32_t log_blowup, uint32_t* out);
/* MerkleTreeMmcs::commit over mixed-height row-major matrices; cap_out 8<<cap_height words. */
int p3r_mmcs_commit(p3r_ctx* ctx, uint32_t n_mats, const p3r_matrix_u32* mats, uint32_t* cap_out);
/* Poseidon2 permutation of n states of 16 words (in place). */
int p3r_poseidon2_permute(p3r_ctx* ctx, uint32_t* states, uint32_t n);

/* Runner assist: the Poseidon2 permutation operations of `CircuitRunner::execute_all`
 * (circuit/src/tables/runner.rs:256-308 -> circuit/src/ops/poseidon_perm/executor.rs:924-975) as chains on the device. Per row, in
 * table order: state = zeros (new_start) or the previous row's output (sponge mode: all four limbs; arity-2 Merkle mode: the
 * first two limbs); on Merkle rows words 8..15 of `values` (the private sibling digest) fill limbs 2, 3; limbs whose bit is set in
 * witness_mask are overwritten with `values` (CTL-exposed witnesses); Merkle rows with mmcs_bit swap the two halves; permute.
 * Rows from a new_start row to the next one form a chain (sequential); chains run in parallel. Writes the resolved input state
 * (what `Poseidon2CircuitRow::input_values` / p3r_poseidon2_ops.input_values holds) and the output state of every row, 16
 * Montgomery words each. D = 4, width 16. At most 32 768 rows per call. */
typedef struct {
    uint32_t n_rows;
    const uint8_t* new_start;      /* row 0 starts a chain whatever its flag says */
    const uint8_t* merkle_path;
    const uint8_t* mmcs_bit;
    const uint8_t* witness_mask;   /* bit l = limb l (words 4l..4l+3) comes from `values` */
    const uint32_t* values;        /* n_rows x 16 Montgomery words */
} p3r_poseidon2_chain_ops;
int p3r_poseidon2_run_chains(p3r_ctx* ctx, const p3r_poseidon2_chain_ops* ops, uint32_t* inputs_out, uint32_t* outputs_out);

/* Host-only Poseidon2 (no CUDA call, usable without a GPU): the permutation the prover's host transcript uses (AVX2 when the
 * CPU has it, checked against the scalar twin at creation), for host code that must hash exactly like the prover — the
 * circuit runner's Poseidon2 rows (circuit/src/ops/poseidon_perm/executor.rs) when a GPU batch is not worth a round trip, and
 * the synthetic-workload generator. `states`: n x 16 Montgomery words, permuted in place. Not a proving fallback. */
typedef struct p3r_host_hasher p3r_host_hasher;
int p3r_host_hasher_create(const p3r_field_desc* field, const p3r_poseidon2_consts* p2, p3r_host_hasher** out);
int p3r_host_hasher_permute(const p3r_host_hasher* h, uint32_t* states, size_t n);
void p3r_host_hasher_free(p3r_host_hasher* h);
/* Device-resident benchmark of the commit path (LDE + Merkle) on a synthetic matrix:
 * returns CUDA-event milliseconds per phase, averaged over `iters`. times_ms_out: [lde, leaves, tree]. */
int p3r_bench_commit(p3r_ctx* ctx, uint32_t log_height, uint32_t width, uint32_t iters, uint64_t seed,
                     float* times_ms_out);

/* Same for a MIXED-HEIGHT batch (heights {2^k, 2^(k-1), 2^(k-3)}): one batched LDE and one MMCS commit over
 * n_mats synthetic matrices. times_ms_out: [lde, commit]. */
int p3r_bench_commit_multi(p3r_ctx* ctx, uint32_t n_mats, const uint32_t* log_heights, const uint32_t* widths, uint32_t iters,
                           uint64_t seed, float* times_ms_out);

/* Device-resident benchmark of one FRI commit round on a synthetic extension-field vector of 2^log_len elements: fold by
 * 2^log_arity (k_fri_fold) and Merkle-commit the folded vector as rows of 2^log_arity elements. times_ms_out: [fold, commit]. */
int p3r_bench_fri_round(p3r_ctx* ctx, uint32_t log_len, uint32_t log_arity, uint32_t iters, uint64_t seed, float* times_ms_out);

/* Per-session device timings (CUDA events on the session stream) of the last one-shot p3r_prove:
 * names_out receives a pointer to a static NUL-separated list; ms_out up to cap entries. */
int p3r_last_phase_times(p3r_ctx* ctx, const char** names_out, float* ms_out, uint32_t cap, uint32_t* n_out);
/* Alternative code paths, for parity tests that compare two independent implementations of the same step. `enable` is a
 * bit set (default 1):
 *   bit 0  set   : build-time specialised (straight-line) quotient kernels (constraint groups) and LogUp-trace kernels (lookup
 *                  groups) for the programs of the recursion layer's ALU / Poseidon2 / Const / Public / Recompose tables
 *                  (scripts/gen_specialized.py) when the program hash matches; clear: always the bytecode interpreter.
 *   bit 1  set   : DISABLE the whole-column LDE kernels (k_ntt_col / k_ntt_top): every LDE goes through the multi-pass tile
 *                  kernel k_ntt_pass.
 *   bit 2  set   : DISABLE the device-side transcript of the FRI commit rounds in p3r_prove* (one host round trip per round).
 *   bit 3  set   : ENABLE the work-queue row hashing (k_hash_rows_queue) instead of one CTA per 64 rows (k_hash_rows); a second
 *                  schedule of the same sponge for the parity tests (measured slower on B200: ptxas gives the persistent loop 40
 *                  registers and a worse pipe mix).
 *   bit 4  set   : DISABLE the constraint-group schedule of the bytecode interpreter (k_quotient_grouped): one thread per row
 *                  runs the whole program (k_quotient). */
int p3r_set_specialization(p3r_ctx* ctx, int enable);

/* ---- proof wire format: flat proof blob <-> postcard bytes of the reference's `BatchStarkProof<SC>`
 * (circuit-prover/src/batch_stark_prover.rs:613-640; `postcard::to_allocvec`, recursion/examples/common/mod.rs:144-147).
 * Host-only (no CUDA call). The field mapping is documented at the top of csrc/wire.cpp. ---- */
typedef struct {                    /* NonPrimitiveTableEntry, batch_stark_prover.rs:274-290 */
    const char* op_type;            /* NpoTypeId string, e.g. "poseidon2_perm/koala_bear_d4_w16", "recompose" */
    uint64_t rows, lanes;
    const uint32_t* public_values;  /* Montgomery words */
    uint32_t n_public_values;
    uint32_t air_variant;           /* AirVariant: 0 Baseline, 1 Optimized (batch_stark_prover.rs:254-260) */
} p3r_npo_entry;
typedef struct {                    /* the BatchStarkProof fields next to `proof` (batch_stark_prover.rs:1598-1641) */
    uint64_t public_lanes, alu_lanes;             /* TablePacking, batch_stark_prover/packing.rs:9-27 */
    const char* const* npo_lane_ops;              /* npo_lanes: Vec<(NpoTypeId, usize)> */
    const uint64_t* npo_lane_counts;
    uint32_t n_npo_lanes;
    uint64_t min_trace_height, horner_packed_steps;
    uint64_t rows[3];                             /* RowCounts: const, public, alu */
    uint32_t alu_variant;
    uint64_t ext_degree;                          /* w_binomial = Some(field.w) iff ext_degree > 1 */
    uint32_t alu_quintic_trinomial;
    const p3r_npo_entry* non_primitives;
    uint32_t n_non_primitives;
    const uint32_t* prep_cap;                     /* stark_common: the preprocessed commitment (p3r_prep_commit's cap), or NULL */
} p3r_proof_meta;
enum {
    P3R_WIRE_CANONICAL = 1,   /* field elements as canonical u32 (default: the Montgomery word, p3's in-memory form) — SURVEY.md B5 */
    P3R_WIRE_BARE_ROOT = 2    /* cap_height 0: commitments as the bare [F; 8] root instead of a one-entry Vec<[F; 8]> cap */
};
/* `descs` are the instances the proof was made for (widths, lookups, quotient chunks); *n_bytes receives the size (also on
 * P3R_ERR_BUFFER); *proof_bytes_out (optional) the length of the leading `proof: BatchProof` field. */
int p3r_proof_serialize(const p3r_field_desc* field, const p3r_fri_params* fri, uint32_t n_inst, const p3r_instance_desc* descs,
                        const uint32_t* blob, size_t n_words, const p3r_proof_meta* meta, uint32_t flags, uint8_t* out,
                        size_t cap_bytes, size_t* n_bytes, size_t* proof_bytes_out);
/* Inverse of the `proof` field: postcard bytes -> flat blob (the metadata that follows is skipped; *proof_bytes = its offset). */
int p3r_proof_deserialize(const p3r_field_desc* field, const p3r_fri_params* fri, const uint8_t* bytes, size_t n_bytes,
                          uint32_t flags, uint32_t* blob_out, size_t cap_words, size_t* n_words, size_t* proof_bytes);
const char* p3r_wire_last_error(void);

/* How host threads wait for the GPU (process-wide): 1 = poll + sched_yield with device->host results staged through pinned
 * memory (default: a waiting thread yields its core to a thread that has kernels to launch; best throughput with several
 * proofs in flight per GPU and few host cores per GPU), 0 = spin inside the driver (cudaStreamSynchronize, direct copies),
 * 2 = block on an event (the thread sleeps; least CPU, highest latency), 3 = poll with a 20 us yielding phase and then 30 us
 * sleeps (for hosts with fewer cores than proving threads, e.g. 8 ranks x 4 proofs in flight on 32 cores). Also settable with the
 * environment variable P3R_WAIT=spin|yield|block|sleep. No reference counterpart (rayon owns the reference's threads). */
void p3r_set_wait_mode(int mode);

/* CUDA stream priority of a context (call between proofs; waits for the context's work first). high != 0: every kernel of the
 * context's proofs is scheduled ahead of pending blocks of default-priority contexts. For servers that mix small,
 * latency-bound proofs (base layers: a few hundred short dependent kernels) with full-size layers on one GPU: without it a
 * small proof's kernels queue behind the thousands of pending blocks of the big proofs' hash launches (measured in the
 * aggregation tree: leaf proofs 1.4 ms among themselves, 2.6 ms next to node proofs). Default: normal priority. */
int p3r_ctx_set_stream_priority(p3r_ctx* ctx, int high);

/* CUDA-event stopwatch on the ctx stream (the stream every kernel of this ctx is launched on): start synchronises the
 * stream and records; stop records, synchronises and returns the elapsed device milliseconds. */
int p3r_timer_start(p3r_ctx* ctx);
int p3r_timer_stop(p3r_ctx* ctx, float* ms_out);

/* Per-kernel-class device timing. p3r_set_kernel_timing enables CUDA-event pairs (on the ctx stream) around every launch
 * group of the classes in `class_mask` (bit k = class k); p3r_kernel_stats returns, per class, the accumulated milliseconds
 * (only for enabled classes), launch counts and ALGORITHMIC bytes (DESIGN.md "Kernels") since the last reset.
 * Classes: 0 ntt_lde, 1 hash_rows, 2 compress, 3 logup, 4 quotient, 5 open, 6 reduced_openings, 7 fri_fold, 8 transpose, 9 misc. */
int p3r_set_kernel_timing(p3r_ctx* ctx, uint32_t class_mask);
int p3r_reset_kernel_stats(p3r_ctx* ctx);
int p3r_kernel_stats(p3r_ctx* ctx, const char** names_out, double* ms_out, uint64_t* launches_out, uint64_t* bytes_out,
                     uint32_t cap, uint32_t* n_out);
/* Poseidon2 permutations issued per class since the last reset (hash_rows / compress): the unit of the INT32-pipe roofline. */
int p3r_kernel_perms(p3r_ctx* ctx, uint64_t* perms_out, uint32_t cap);
/* Number of kernel launches issued by this ctx since creation (for bench.py's gpu_launches). */
uint64_t p3r_launch_count(const p3r_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* P3R_H */
