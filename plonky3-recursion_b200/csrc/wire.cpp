// wire.cpp — proof wire format: flat proof blob (DESIGN.md "Proof blob")  <->  postcard bytes of the reference's
// `BatchStarkProof<SC>`. Host-only code (no CUDA call), part of libp3r_b200.so.
//
// What is encoded, in serde field order:
//   BatchStarkProof { proof, table_packing, rows, alu_variant, ext_degree, w_binomial, alu_quintic_trinomial, non_primitives,
//                     stark_common }                                  circuit-prover/src/batch_stark_prover.rs:613-640
//   proof: BatchProof { commitments, opened_values, opening_proof, lookup_terminals, degree_bits }
//                                                                     recursion/src/generation.rs:95-101
//     commitments { main, permutation: Option, quotient_chunks, random: Option }          recursion/src/types/proof.rs:258-263
//     opened_values.instances[i] { base_opened_values { trace_local, trace_next: Option, preprocessed_local: Option,
//                                  preprocessed_next: Option, quotient_chunks: Vec<Vec<EF>>, random: Option },
//                                  permutation_local, permutation_next }  recursion/src/generation.rs:278-287,357-360;
//                                                                         recursion/src/types/proof.rs:82-125
//     opening_proof: FriProof { commit_phase_commits, commit_pow_witnesses, query_proofs, final_poly, query_pow_witness }
//                                                                     recursion/src/pcs/fri/targets.rs:42-45,105-109
//       query_proofs[q] { input_proof: Vec<BatchOpening { opened_values: Vec<Vec<F>>, opening_proof: Vec<[F; 8]> }>,
//                         commit_phase_openings: Vec<{ log_arity: u8, sibling_values: Vec<EF>, opening_proof }> }
//                                                                     recursion/src/pcs/fri/targets.rs:146-147,212-216,307-309
//   table_packing: TablePacking { public_lanes, alu_lanes, npo_lanes: Vec<(NpoTypeId, usize)>, min_trace_height,
//                                 horner_packed_steps }               circuit-prover/src/batch_stark_prover/packing.rs:9-27
//   rows: RowCounts([usize; 3])                                       batch_stark_prover.rs:460
//   alu_variant / air_variant: AirVariant (unit enum, variant index)  batch_stark_prover.rs:254-260
//   non_primitives[k] { op_type: NpoTypeId(String), rows, lanes, public_values: Vec<F>, air_variant }   :274-290
//   stark_common: Option<SerializedStarkCommon { commitment, instances: Vec<Option<{matrix_index, width, degree_bits}>>,
//                                                matrix_to_instance }>      batch_stark_prover.rs:492-512,582-600
// postcard (the format `report_proof_size` uses, recursion/examples/common/mod.rs:144-147): unsigned integers wider than a byte
// are LEB128 varints, u8 / bool one byte, Option = 0x00 | 0x01 + value, Vec / String = varint length + items, tuples, arrays
// and structs = their fields back to back, unit enum variant = varint index.
// [P3-EXT] (crates.io p3-* 0.6, not in the reference tree; selectable with `flags`):
//   * a field element is `serialize_u32` of its Montgomery word (P3R_WIRE_CANONICAL: of its canonical value) — SURVEY.md B5;
//   * an extension element is the tuple of its 4 coefficients; a digest is the array [F; 8];
//   * a commitment is the Merkle cap `Vec<[F; 8]>` (P3R_WIRE_BARE_ROOT with cap_height 0: the bare `[F; 8]` root).
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/p3r.h"

namespace {

struct Field {
    uint32_t p, mu_neg;   // mu_neg = -p^-1 mod 2^32
    uint32_t r2;          // 2^64 mod p
    explicit Field(uint32_t p_) : p(p_) {
        uint32_t inv = p;                       // Newton: p * inv = 1 mod 2^32
        for (int i = 0; i < 5; i++) inv *= 2u - p * inv;
        mu_neg = 0u - inv;
        uint64_t r = (((uint64_t)1 << 32) % p);
        r2 = (uint32_t)((r * r) % p);
    }
    uint32_t mont_mul(uint32_t a, uint32_t b) const {
        uint64_t t = (uint64_t)a * b;
        uint32_t m = (uint32_t)t * mu_neg;
        uint64_t u = (t + (uint64_t)m * p) >> 32;
        return (uint32_t)(u >= p ? u - p : u);
    }
    uint32_t from_monty(uint32_t m) const { return mont_mul(m, 1u); }
    uint32_t to_monty(uint32_t c) const { return mont_mul(c % p, r2); }
};

struct Writer {
    std::vector<uint8_t> b;
    const Field& f;
    bool canonical;
    Writer(const Field& f_, bool c) : f(f_), canonical(c) {}
    void u8(uint8_t v) { b.push_back(v); }
    void varint(uint64_t v) {
        while (v >= 0x80) {
            b.push_back((uint8_t)(v | 0x80));
            v >>= 7;
        }
        b.push_back((uint8_t)v);
    }
    void fe(uint32_t monty) { varint(canonical ? f.from_monty(monty) : monty); }
    void fes(const uint32_t* w, size_t n) {
        for (size_t i = 0; i < n; i++) fe(w[i]);
    }
    void vec_ef(const uint32_t* w, size_t n_ef) {   // Vec<EF>
        varint(n_ef);
        fes(w, 4 * n_ef);
    }
    void str(const char* s) {
        size_t n = s ? std::strlen(s) : 0;
        varint(n);
        b.insert(b.end(), s, s + n);
    }
};

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    const Field& f;
    bool canonical;
    bool ok = true;
    Reader(const uint8_t* b, size_t n, const Field& f_, bool c) : p(b), end(b + n), f(f_), canonical(c) {}
    uint8_t u8() {
        if (p >= end) {
            ok = false;
            return 0;
        }
        return *p++;
    }
    uint64_t varint() {
        uint64_t v = 0;
        for (int shift = 0; shift < 64; shift += 7) {
            uint8_t c = u8();
            v |= (uint64_t)(c & 0x7f) << shift;
            if (!(c & 0x80)) return v;
        }
        ok = false;
        return v;
    }
    uint32_t fe() {
        uint64_t v = varint();
        if (v >= f.p) ok = false;
        return canonical ? f.to_monty((uint32_t)v) : (uint32_t)v;
    }
    size_t len(size_t max = (size_t)1 << 28) {
        uint64_t n = varint();
        if (n > max) {
            ok = false;
            return 0;
        }
        return (size_t)n;
    }
};

thread_local std::string g_wire_err;
int fail(const std::string& s, int code = P3R_ERR_INVALID_ARG) {
    g_wire_err = s;
    return code;
}

// Sequential cursor over the blob.
struct Blob {
    const uint32_t* w;
    size_t n, pos = 0;
    bool ok = true;
    const uint32_t* take(size_t k) {
        if (pos + k > n) {   // truncated: flag it and hand out zeros so the caller can finish its statement safely
            ok = false;
            pos = n;
            static thread_local std::vector<uint32_t> zeros;
            if (zeros.size() < k) zeros.assign(k, 0);
            return zeros.data();
        }
        const uint32_t* r = w + pos;
        pos += k;
        return r;
    }
    uint32_t one() {
        const uint32_t* r = take(1);
        return r ? *r : 0;
    }
};

void write_commitment(Writer& w, const uint32_t* cap, uint32_t cap_words, bool bare_root) {
    if (bare_root && cap_words == 8) {
        w.fes(cap, 8);
        return;
    }
    w.varint(cap_words / 8);
    w.fes(cap, cap_words);
}
void write_path(Writer& w, const uint32_t* digests, size_t depth) {   // Vec<[F; 8]>
    w.varint(depth);
    w.fes(digests, depth * 8);
}

}  // namespace

extern "C" {

const char* p3r_wire_last_error(void) { return g_wire_err.c_str(); }

int p3r_proof_serialize(const p3r_field_desc* field, const p3r_fri_params* fri, uint32_t n_inst, const p3r_instance_desc* descs,
                        const uint32_t* blob, size_t n_words, const p3r_proof_meta* meta, uint32_t flags, uint8_t* out,
                        size_t cap_bytes, size_t* n_bytes, size_t* proof_bytes_out) {
    if (!field || !fri || !descs || !blob || !meta || !n_bytes || n_inst == 0) return fail("proof_serialize: null argument");
    const Field F(field->p);
    Writer w(F, (flags & P3R_WIRE_CANONICAL) != 0);
    const bool bare_root = (flags & P3R_WIRE_BARE_ROOT) != 0 && fri->cap_height == 0;
    Blob b{blob, n_words};
    const uint32_t* hdr = b.take(5);
    if (!b.ok || hdr[0] != 0x50335250u || hdr[1] != n_inst) return fail("proof_serialize: not a proof blob of this instance count");
    const bool has_perm = hdr[2] != 0, has_prep = hdr[3] != 0;
    const uint32_t capw = hdr[4];
    if (capw != (8u << fri->cap_height)) return fail("proof_serialize: cap size does not match cap_height");
    const uint32_t* degree_bits = b.take(n_inst);
    const uint32_t* main_cap = b.take(capw);
    const uint32_t* perm_cap = has_perm ? b.take(capw) : nullptr;
    const uint32_t* quot_cap = b.take(capw);
    size_t n_perm_inst = 0;
    uint32_t log_max = 0;
    for (uint32_t i = 0; i < n_inst; i++) {
        n_perm_inst += descs[i].n_lookups != 0;
        if (!b.ok || degree_bits[i] != descs[i].log_height) return fail("proof_serialize: degree_bits differ from the instance list");
        log_max = std::max(log_max, descs[i].log_height + fri->log_blowup);
    }
    const uint32_t* terminals = has_perm ? b.take(4 * n_perm_inst) : nullptr;
    if (!b.ok) return fail("proof_serialize: truncated blob");
    // ---- BatchProof.commitments ----
    write_commitment(w, main_cap, capw, bare_root);
    w.u8(has_perm ? 1 : 0);
    if (has_perm) write_commitment(w, perm_cap, capw, bare_root);
    write_commitment(w, quot_cap, capw, bare_root);
    w.u8(0);   // random: None (non-ZK)
    // ---- opened_values.instances ----
    w.varint(n_inst);
    for (uint32_t i = 0; i < n_inst; i++) {
        const p3r_instance_desc& d = descs[i];
        const uint32_t aux = d.n_lookups ? d.n_lookups + 1 : 0, n_chunks = 1u << d.log_quotient_chunks;
        w.vec_ef(b.take(4 * (size_t)d.main_width), d.main_width);                       // trace_local
        w.u8(d.uses_next_row ? 1 : 0);                                                   // trace_next
        if (d.uses_next_row) w.vec_ef(b.take(4 * (size_t)d.main_width), d.main_width);
        w.u8(d.prep_width ? 1 : 0);                                                      // preprocessed_local
        if (d.prep_width) w.vec_ef(b.take(4 * (size_t)d.prep_width), d.prep_width);
        w.u8(d.prep_width ? 1 : 0);                                                      // preprocessed_next
        if (d.prep_width) w.vec_ef(b.take(4 * (size_t)d.prep_width), d.prep_width);
        const uint32_t* perm_local = b.take(16 * (size_t)aux);
        const uint32_t* perm_next = b.take(16 * (size_t)aux);
        w.varint(n_chunks);                                                              // quotient_chunks
        for (uint32_t c = 0; c < n_chunks; c++) w.vec_ef(b.take(16), 4);
        w.u8(0);                                                                         // random: None
        w.vec_ef(perm_local, 4 * (size_t)aux);                                           // permutation_local
        w.vec_ef(perm_next, 4 * (size_t)aux);                                            // permutation_next
        if (!b.ok) return fail("proof_serialize: truncated opened values");
    }
    // ---- opening_proof: FriProof ----
    const uint32_t n_rounds = b.one();
    if (!b.ok || n_rounds > 32) return fail("proof_serialize: bad FRI round count");
    const uint32_t* log_arities = b.take(n_rounds);
    const uint32_t* fri_caps = b.take((size_t)n_rounds * capw);
    const uint32_t* commit_pow = b.take(n_rounds);
    const size_t n_final = (size_t)1 << fri->log_final_poly_len;
    const uint32_t* final_poly = b.take(4 * n_final);
    const uint32_t query_pow = b.one();
    if (!b.ok) return fail("proof_serialize: truncated FRI header");
    w.varint(n_rounds);
    for (uint32_t r = 0; r < n_rounds; r++) write_commitment(w, fri_caps + (size_t)r * capw, capw, bare_root);
    w.varint(n_rounds);
    w.fes(commit_pow, n_rounds);
    // input rounds [main, quotient, preprocessed?, permutation?]: matrices in commit order
    struct MatW {
        uint32_t width, log_h;
    };
    std::vector<std::vector<MatW>> rounds(4);
    for (uint32_t i = 0; i < n_inst; i++) {
        const p3r_instance_desc& d = descs[i];
        const uint32_t lh = d.log_height + fri->log_blowup;
        rounds[0].push_back({d.main_width, lh});
        for (uint32_t c = 0; c < (1u << d.log_quotient_chunks); c++) rounds[1].push_back({4, lh});
        if (d.prep_width) rounds[2].push_back({d.prep_width, lh});
        if (d.n_lookups) rounds[3].push_back({(d.n_lookups + 1) * 4, lh});
    }
    std::vector<int> order = {0, 1};
    if (has_prep) order.push_back(2);
    if (has_perm) order.push_back(3);
    w.varint(fri->num_queries);
    for (uint32_t q = 0; q < fri->num_queries; q++) {
        w.varint(order.size());                                    // input_proof: one BatchOpening per input round
        for (int k : order) {
            uint32_t lmax = 0;
            w.varint(rounds[k].size());                            // opened_values: Vec<Vec<F>>
            for (const MatW& m : rounds[k]) {
                w.varint(m.width);
                w.fes(b.take(m.width), m.width);
                lmax = std::max(lmax, m.log_h);
            }
            const size_t depth = lmax - fri->cap_height;
            write_path(w, b.take(depth * 8), depth);
        }
        w.varint(n_rounds);                                        // commit_phase_openings
        uint32_t h = log_max;
        for (uint32_t r = 0; r < n_rounds; r++) {
            const uint32_t la = log_arities[r], arity = 1u << la;
            if (la == 0 || la > 8 || la > h) return fail("proof_serialize: bad log_arity");
            w.u8((uint8_t)la);
            w.vec_ef(b.take(4 * (size_t)(arity - 1)), arity - 1);
            h -= la;
            const size_t depth = h - fri->cap_height;
            write_path(w, b.take(depth * 8), depth);
        }
        if (!b.ok) return fail("proof_serialize: truncated query proof");
    }
    w.vec_ef(final_poly, n_final);
    w.fe(query_pow);
    if (b.pos != n_words) return fail("proof_serialize: trailing words in the blob");
    // ---- lookup_terminals: Vec<Option<EF>>, degree_bits: Vec<usize> ----
    w.varint(n_inst);
    size_t tk = 0;
    for (uint32_t i = 0; i < n_inst; i++) {
        const bool has = descs[i].n_lookups != 0;
        w.u8(has ? 1 : 0);
        if (has) w.fes(terminals + 4 * (tk++), 4);
    }
    w.varint(n_inst);
    for (uint32_t i = 0; i < n_inst; i++) w.varint(degree_bits[i]);
    const size_t proof_bytes = w.b.size();
    // ---- metadata (batch_stark_prover.rs:1598-1641) ----
    w.varint(meta->public_lanes);
    w.varint(meta->alu_lanes);
    w.varint(meta->n_npo_lanes);
    for (uint32_t k = 0; k < meta->n_npo_lanes; k++) {
        w.str(meta->npo_lane_ops[k]);
        w.varint(meta->npo_lane_counts[k]);
    }
    w.varint(meta->min_trace_height);
    w.varint(meta->horner_packed_steps);
    for (int k = 0; k < 3; k++) w.varint(meta->rows[k]);
    w.varint(meta->alu_variant);
    w.varint(meta->ext_degree);
    w.u8(meta->ext_degree > 1 ? 1 : 0);
    if (meta->ext_degree > 1) w.fe(F.to_monty(field->w));
    w.u8(meta->alu_quintic_trinomial ? 1 : 0);
    w.varint(meta->n_non_primitives);
    for (uint32_t k = 0; k < meta->n_non_primitives; k++) {
        const p3r_npo_entry& e = meta->non_primitives[k];
        w.str(e.op_type);
        w.varint(e.rows);
        w.varint(e.lanes);
        w.varint(e.n_public_values);
        w.fes(e.public_values, e.n_public_values);
        w.varint(e.air_variant);
    }
    w.u8(has_prep ? 1 : 0);                                        // stark_common: Option<SerializedStarkCommon>
    if (has_prep) {
        if (!meta->prep_cap) return fail("proof_serialize: the proof has preprocessed columns but meta->prep_cap is NULL");
        write_commitment(w, meta->prep_cap, capw, bare_root);
        w.varint(n_inst);
        std::vector<uint32_t> m2i;
        for (uint32_t i = 0; i < n_inst; i++) {
            w.u8(descs[i].prep_width ? 1 : 0);
            if (descs[i].prep_width) {
                w.varint(m2i.size());
                w.varint(descs[i].prep_width);
                w.varint(descs[i].log_height);
                m2i.push_back(i);
            }
        }
        w.varint(m2i.size());
        for (uint32_t i : m2i) w.varint(i);
    }
    *n_bytes = w.b.size();
    if (proof_bytes_out) *proof_bytes_out = proof_bytes;
    if (w.b.size() > cap_bytes || !out) return fail("proof_serialize: output buffer too small", P3R_ERR_BUFFER);
    std::memcpy(out, w.b.data(), w.b.size());
    return P3R_OK;
}

int p3r_proof_deserialize(const p3r_field_desc* field, const p3r_fri_params* fri, const uint8_t* bytes, size_t n_bytes,
                          uint32_t flags, uint32_t* blob_out, size_t cap_words, size_t* n_words, size_t* proof_bytes) {
    if (!field || !fri || !bytes || !n_words) return fail("proof_deserialize: null argument");
    const Field F(field->p);
    Reader r(bytes, n_bytes, F, (flags & P3R_WIRE_CANONICAL) != 0);
    const bool bare_root = (flags & P3R_WIRE_BARE_ROOT) != 0 && fri->cap_height == 0;
    const uint32_t capw = 8u << fri->cap_height;
    auto read_commitment = [&](std::vector<uint32_t>& dst) {
        if (!bare_root && r.len() != capw / 8) r.ok = false;
        for (uint32_t k = 0; k < capw; k++) dst.push_back(r.fe());
    };
    auto read_vec_ef = [&](std::vector<uint32_t>& dst) {
        size_t n = r.len();
        for (size_t k = 0; k < 4 * n && r.ok; k++) dst.push_back(r.fe());
        return n;
    };
    std::vector<uint32_t> main_cap, perm_cap, quot_cap, opened, fri_part, queries, terminals, degree_bits;
    read_commitment(main_cap);
    const bool has_perm = r.u8() != 0;
    if (has_perm) read_commitment(perm_cap);
    read_commitment(quot_cap);
    if (r.u8() != 0) return fail("proof_deserialize: ZK proofs (random commitment) are not supported");
    const size_t n_inst = r.len(1 << 16);
    bool has_prep = false;
    for (size_t i = 0; i < n_inst && r.ok; i++) {
        std::vector<uint32_t> perm_l, perm_n, chunks;
        read_vec_ef(opened);
        if (r.u8()) read_vec_ef(opened);
        for (int k = 0; k < 2; k++)
            if (r.u8()) {
                has_prep = true;
                read_vec_ef(opened);
            }
        const size_t n_chunks = r.len(1 << 8);
        for (size_t c = 0; c < n_chunks; c++)
            if (read_vec_ef(chunks) != 4) r.ok = false;
        if (r.u8() != 0) return fail("proof_deserialize: ZK proofs (random opened values) are not supported");
        read_vec_ef(perm_l);
        read_vec_ef(perm_n);
        opened.insert(opened.end(), perm_l.begin(), perm_l.end());
        opened.insert(opened.end(), perm_n.begin(), perm_n.end());
        opened.insert(opened.end(), chunks.begin(), chunks.end());
    }
    const size_t n_rounds = r.len(32);
    std::vector<uint32_t> fri_caps, commit_pow, log_arities, final_poly;
    for (size_t k = 0; k < n_rounds; k++) read_commitment(fri_caps);
    if (r.len(32) != n_rounds) r.ok = false;
    for (size_t k = 0; k < n_rounds; k++) commit_pow.push_back(r.fe());
    const size_t nq = r.len(1 << 12);
    for (size_t q = 0; q < nq && r.ok; q++) {
        const size_t n_in = r.len(8);
        for (size_t k = 0; k < n_in && r.ok; k++) {
            const size_t n_mats = r.len(1 << 16);
            for (size_t m = 0; m < n_mats && r.ok; m++) {
                const size_t wd = r.len(1 << 20);
                for (size_t c = 0; c < wd && r.ok; c++) queries.push_back(r.fe());
            }
            const size_t depth = r.len(64);
            for (size_t c = 0; c < depth * 8 && r.ok; c++) queries.push_back(r.fe());
        }
        if (r.len(32) != n_rounds) r.ok = false;
        for (size_t k = 0; k < n_rounds && r.ok; k++) {
            const uint32_t la = r.u8();
            if (q == 0) log_arities.push_back(la);
            else if (log_arities[k] != la) r.ok = false;
            if (read_vec_ef(queries) != ((size_t)1 << la) - 1) r.ok = false;
            const size_t depth = r.len(64);
            for (size_t c = 0; c < depth * 8 && r.ok; c++) queries.push_back(r.fe());
        }
    }
    read_vec_ef(final_poly);
    const uint32_t query_pow = r.fe();
    if (r.len(1 << 16) != n_inst) r.ok = false;
    for (size_t i = 0; i < n_inst && r.ok; i++)
        if (r.u8())
            for (int k = 0; k < 4; k++) terminals.push_back(r.fe());
    if (r.len(1 << 16) != n_inst) r.ok = false;
    for (size_t i = 0; i < n_inst && r.ok; i++) degree_bits.push_back((uint32_t)r.varint());
    if (!r.ok) return fail("proof_deserialize: malformed or truncated bytes");
    if (proof_bytes) *proof_bytes = (size_t)(r.p - bytes);
    std::vector<uint32_t> blob = {0x50335250u, (uint32_t)n_inst, has_perm ? 1u : 0u, has_prep ? 1u : 0u, capw};
    auto put = [&](const std::vector<uint32_t>& v) { blob.insert(blob.end(), v.begin(), v.end()); };
    put(degree_bits);
    put(main_cap);
    put(perm_cap);
    put(quot_cap);
    put(terminals);
    put(opened);
    blob.push_back((uint32_t)n_rounds);
    put(log_arities);
    put(fri_caps);
    put(commit_pow);
    put(final_poly);
    blob.push_back(query_pow);
    put(queries);
    *n_words = blob.size();
    if (blob.size() > cap_words || !blob_out) return fail("proof_deserialize: output buffer too small", P3R_ERR_BUFFER);
    std::memcpy(blob_out, blob.data(), blob.size() * 4);
    return P3R_OK;
}

}  // extern "C"
