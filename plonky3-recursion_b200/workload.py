"""Synthetic recursion-layer workloads (the bench / test inputs).

The reference's workload is the verifier circuit of the previous proof, interpreted by the host `CircuitRunner`
(/root/reference circuit/src/tables/runner.rs:194-249, out of scope here: it stays host Rust). What reaches the hot path
is `Traces<EF>`: lists of Const / Public / ALU / Poseidon2 / Recompose operations over a shared witness bus. This module
generates such operation lists directly — random but VALID (every ALU relation holds, every Poseidon2 row is a real
permutation, the WitnessChecks bus balances) — at the table shapes of the reference's steady-state recursion layer
(ALU 3 lanes x D4 x k=4 -> 80 main / 60 preprocessed columns, Poseidon2 KoalaBear 166 (+mmcs), ...),
then turns them into matrices with the same builders the reference uses (`trace_to_matrix`, preprocessed columns with
multiplicities, circuit-prover/src/common.rs:127-390).
"""
from __future__ import annotations

import numpy as np

from . import air
from .airs import alu, poseidon2, witness_send
from .poseidon2_params import Poseidon2Params
from .field import Field

D = 4


class Witnesses:
    """Witness bus bookkeeping: values (EF4) and read counts (ext_reads, common.rs:226-287)."""

    def __init__(self):
        self.values = []
        self.reads = []

    def new(self, v):
        self.values.append([int(x) for x in v])
        self.reads.append(0)
        return len(self.values) - 1

    def read(self, wid):
        self.reads[wid] += 1
        return self.values[wid]


def _rand_ext(F: Field, rng):
    return [int(x) for x in rng.integers(0, F.p, size=D)]


def gen_alu_ops(F: Field, rng, W: Witnesses, sources, n_ops: int, horner_frac=0.35, chain_len=(6, 24)):
    """Random valid ALU op list. `sources`: witness ids that already exist (Const/Public created).
    Returns alu.AluOps plus the list of created witness ids. Creator multiplicities are patched in `finalize_alu`."""
    p = F.p
    vals, preps, created_out = [], [], []
    pool = list(sources)

    def pick():
        return pool[int(rng.integers(0, len(pool)))] if rng.random() < 0.5 else pool[-1 - int(rng.integers(0, min(len(pool), 64)))]

    def emul(x, y):
        return F.ext_mul(x, y)

    i = 0
    while i < n_ops:
        if rng.random() < horner_frac and n_ops - i >= 4:
            L = int(rng.integers(chain_len[0], chain_len[1] + 1))
            L = min(L, n_ops - i)
            b_id = pick()
            acc = [0] * D
            for t in range(L):
                a_id, c_id = pick(), pick()
                a, c, bv = W.read(a_id), W.read(c_id), W.read(b_id)
                acc = [(u + v - w_) % p for u, v, w_ in zip(emul(acc, bv), c, a)]
                out_id = W.new(acc)
                created_out.append(out_id)
                vals.append([a, bv, c, acc])
                pr = [0] * 13
                pr[alu.MULT_A] = p - 1
                pr[alu.SEL_HORNER] = 1
                pr[alu.A_IDX], pr[alu.B_IDX], pr[alu.C_IDX], pr[alu.OUT_IDX] = a_id * D, b_id * D, c_id * D, out_id * D
                pr[alu.MULT_B] = p - 1
                pr[alu.MULT_OUT] = -1  # creator: patched with n_reads
                pr[alu.A_READER], pr[alu.C_READER] = 1, 1
                preps.append(pr)
                if t == L - 1:
                    pool.append(out_id)  # only the chain's final accumulator is visible to later ops
            i += L
            # a non-Horner op must separate two chains (maximal runs are one chain in compute_schedule)
            if i >= n_ops:
                break
        kind = [alu.ADD, alu.MUL, alu.MULADD, alu.BOOL][int(rng.choice(4, p=[0.3, 0.4, 0.25, 0.05]))]
        a_id, b_id = pick(), pick()
        pr = [0] * 13
        pr[alu.MULT_A] = p - 1
        pr[alu.MULT_B] = p - 1
        pr[alu.MULT_OUT] = -1
        pr[alu.A_READER] = 1
        c = [0] * D
        if kind == alu.BOOL:
            bit = int(rng.integers(0, 2))
            a_id = W.new([bit, 0, 0, 0])  # a private boolean: created by this row (a_state = 2 -> A_READER = -n_reads)
            pr[alu.A_READER] = -2
            a = W.values[a_id]
            bv = W.read(b_id)
            out = list(a)
            pr[alu.SEL_BOOL] = 1
            pool.append(a_id)
        else:
            a, bv = W.read(a_id), W.read(b_id)
            if kind == alu.ADD:
                out = [(u + v) % p for u, v in zip(a, bv)]
                pr[alu.SEL_ADD] = 1
            elif kind == alu.MUL:
                out = emul(a, bv)
            else:
                c_id = pick()
                c = W.read(c_id)
                out = [(u + v) % p for u, v in zip(emul(a, bv), c)]
                pr[alu.SEL_MULADD] = 1
                pr[alu.C_IDX] = c_id * D
                pr[alu.C_READER] = 1
        out_id = W.new(out)
        created_out.append(out_id)
        pool.append(out_id)
        pr[alu.A_IDX], pr[alu.B_IDX], pr[alu.OUT_IDX] = a_id * D, b_id * D, out_id * D
        vals.append([a, bv, c, out])
        preps.append(pr)
        i += 1
    return vals, preps, pool


def finalize_alu(F: Field, W: Witnesses, vals, preps) -> alu.AluOps:
    """Patch creator multiplicities with the final read counts (common.rs:236-268)."""
    p = F.p
    arr = np.zeros((len(preps), 13), dtype=np.uint32)
    for i, pr in enumerate(preps):
        pr = list(pr)
        if pr[alu.MULT_OUT] == -1:
            pr[alu.MULT_OUT] = W.reads[pr[alu.OUT_IDX] // D] % p
        if pr[alu.A_READER] == -2:  # private creator: column = -(n_reads) so that mult_a * col = +n_reads
            pr[alu.A_READER] = (-W.reads[pr[alu.A_IDX] // D]) % p
        arr[i] = [int(x) % p for x in pr]
    return alu.AluOps(np.array(vals, dtype=np.uint32).reshape(-1, 4, D), arr)


class NpoTables:
    """Poseidon2 + Recompose operation lists of a synthetic layer (NPO registration order [Poseidon2, Recompose],
    recursion/src/backend/fri.rs:693-721). Operations read Const/Public witnesses and create new ones that the ALU reads."""

    def __init__(self, F: Field, rng, W: Witnesses, sources, new_public, n_perms: int, n_recompose: int,
                 recompose_coeff: bool = False):
        self.F, self.params = F, Poseidon2Params(F.field_id)
        p = F.p
        rows = []  # dicts, one per permutation row
        created = []

        def pick():
            return sources[int(rng.integers(0, len(sources)))]

        # One dependent permutation per table row: the library's host hasher (p3r_host_hasher, 20 us per call) when it is
        # built, the numpy restatement (1 ms per call, same values) otherwise — this is input generation, not the prover.
        hasher = None
        try:
            from . import lib as _lib
            hasher = _lib.HostHasher(F, self.params)
        except Exception:
            hasher = None

        def perm(state):
            if hasher is not None:
                return [int(x) for x in hasher.permute(np.array(state, dtype=np.uint32))[0]]
            return [int(x) for x in self.params.permute(np.array(state, dtype=np.uint64).reshape(1, 16))[0]]

        def row(**kw):
            r = dict(new_start=False, merkle_path=False, mmcs_bit=False, mmcs_index_sum=0, inputs=None, in_ctl=[False] * 4,
                     in_idx=[0] * 4, out_ids=[None, None], mmcs_idx=0, mmcs_en=False)
            r.update(kw)
            rows.append(r)
            return r

        while len(rows) < n_perms:
            remaining = n_perms - len(rows)
            if rng.random() < 0.5 and remaining >= 12:
                # Merkle chain: one leaf-hash sponge row then M directional compress rows, root exposed, index sent
                M = int(rng.integers(8, 18))
                M = min(M, remaining - 1)
                ids = [pick() for _ in range(4)]
                state = sum((W.read(i) for i in ids), [])
                out = perm(state)
                row(new_start=True, inputs=state, in_ctl=[True] * 4, in_idx=ids)
                acc = 0
                for t in range(M):
                    bit = bool(rng.integers(0, 2))
                    sib = [int(x) for x in rng.integers(0, p, size=8)]
                    state = (sib + out[:8]) if bit else (out[:8] + sib)
                    acc = (acc * 2 + int(bit)) % p
                    r = row(merkle_path=True, mmcs_bit=bit, inputs=state)
                    out = perm(state)
                    if t == M - 1:
                        idx_w = new_public([acc, 0, 0, 0])
                        W.read(idx_w)
                        r["mmcs_idx"], r["mmcs_en"] = idx_w, True
                        r["out_ids"] = [W.new(out[0:4]), W.new(out[4:8])]
                        created.extend(r["out_ids"])
            else:
                # sponge chain
                Ls = int(rng.integers(1, 7))
                Ls = min(Ls, remaining)
                out = None
                for t in range(Ls):
                    if t == 0:
                        ids = [pick() for _ in range(4)]
                        state = sum((W.read(i) for i in ids), [])
                        r = row(new_start=True, inputs=state, in_ctl=[True] * 4, in_idx=ids)
                    else:
                        ids = [pick(), pick()]
                        state = W.read(ids[0]) + W.read(ids[1]) + out[8:16]
                        r = row(inputs=state, in_ctl=[True, True, False, False], in_idx=ids + [0, 0])
                    out = perm(state)
                    if t == Ls - 1:
                        r["out_ids"] = [W.new(out[0:4]), W.new(out[4:8])]
                        created.extend(r["out_ids"])
        self.rows = rows
        # Recompose: EF witnesses packed from base coefficients (creator side only)
        self.recompose_ids = [W.new(_rand_ext(F, rng)) for _ in range(n_recompose)]
        self.created = created + self.recompose_ids
        # `recompose/coeff` (recompose_air.rs:175-197): every coefficient v_i is itself a witness (v_i, 0, 0, 0). Even
        # operations create theirs (hint outputs: multiplicity = number of reads); odd operations point at witnesses a
        # Public row creates, so their coefficient multiplicity is 0 (batch_stark_prover/recompose.rs:341-352).
        self.recompose_coeff = recompose_coeff
        self.coeff_ids, self.coeff_hint = [], []
        if recompose_coeff:
            for k, rid in enumerate(self.recompose_ids):
                hint = k % 2 == 0
                ids = [(W.new if hint else new_public)([c, 0, 0, 0]) for c in W.values[rid]]
                self.coeff_ids.append(ids)
                self.coeff_hint.append(hint)
                self.created.extend(ids)

    def finish(self, W: Witnesses, buses, min_height: int, recompose_lanes: int = 1):
        F, params = self.F, self.params
        ops = poseidon2.Poseidon2Ops(len(self.rows))
        for i, r in enumerate(self.rows):
            ops.new_start[i], ops.merkle_path[i], ops.mmcs_bit[i] = r["new_start"], r["merkle_path"], r["mmcs_bit"]
            ops.input_values[i] = r["inputs"]
            ops.in_ctl[i] = r["in_ctl"]
            ops.input_indices[i] = r["in_idx"]
            for k in range(2):
                if r["out_ids"][k] is not None:
                    ops.output_indices[i, k] = r["out_ids"][k]
                    ops.out_mult[i, k] = W.reads[r["out_ids"][k]]
            ops.mmcs_index_sum_idx[i] = r["mmcs_idx"]
            ops.mmcs_ctl_enabled[i] = r["mmcs_en"]
        self.ops = ops
        tp2, pp2 = poseidon2.build_tables(params, ops, min_height)
        lh = lambda m: int(m.shape[0]).bit_length() - 1
        w2, pw2 = poseidon2.widths(params)
        insts = [air.build_instance("poseidon2", poseidon2.make_eval(params), F.p, lh(tp2), w2, pw2, 0, buses)]
        preps, traces = [pp2], [tp2]
        if self.recompose_ids:
            v = np.array([W.values[i] for i in self.recompose_ids], dtype=np.uint32).reshape(-1, D)
            m = np.array([W.reads[i] for i in self.recompose_ids], dtype=np.uint32)
            idx = np.array(self.recompose_ids, dtype=np.uint32) * D
            tr = witness_send.trace_to_matrix(v, D, recompose_lanes, min_height)
            cidx = cmult = None
            if self.recompose_coeff:
                cidx = np.array(self.coeff_ids, dtype=np.uint32) * D
                cmult = np.array([[W.reads[c] if h else 0 for c in ids] for ids, h in zip(self.coeff_ids, self.coeff_hint)],
                                 dtype=np.uint32)
            pr = witness_send.preprocessed_matrix(m, idx, recompose_lanes, min_height, idx_first=True,
                                                  coeff_idxs=cidx, coeff_mults=cmult)
            plw = witness_send.prep_lane_width(D, self.recompose_coeff)
            insts.append(air.build_instance("recompose/coeff" if self.recompose_coeff else "recompose",
                                            witness_send.make_eval(D, recompose_lanes, idx_first=True,
                                                                   coeff_lookups=self.recompose_coeff),
                                            F.p, lh(tr), D * recompose_lanes, plw * recompose_lanes, 0, buses))
            preps.append(pr)
            traces.append(tr)
        return insts, preps, traces, [None] * len(insts)


class LayerWorkload:
    def __init__(self, insts, preps, traces, pubs, shapes, p2_ops=None, alu_ops=None):
        self.insts, self.preps, self.traces, self.pubs, self.shapes = insts, preps, traces, pubs, shapes
        self.p2_ops = p2_ops or {}   # {instance index: Poseidon2Ops} for the GPU table-fill path
        self.alu_ops = alu_ops or {}  # {instance index: airs.alu.AluTableOps}

    @property
    def h2d_bytes(self):
        return int(sum(t.size * 4 for t in self.traces))


def synthetic_layer(F: Field, seed: int, n_const: int, n_public: int, n_alu: int, n_perms: int = 0, n_recompose: int = 0,
                    alu_lanes: int = 3, horner_k: int = 4, public_lanes: int = 1, recompose_lanes: int = 1,
                    min_height: int = 256, recompose_coeff: bool = False) -> LayerWorkload:
    """Tables in the reference's instance order [Const, Public, ALU, Poseidon2, Recompose]
    (circuit-prover/src/batch_stark_prover.rs:1493-1519; NPO order recursion/src/backend/fri.rs:693-721)."""
    rng = np.random.default_rng(seed)
    W = Witnesses()
    buses = air.BusRegistry()
    const_ids = [W.new(_rand_ext(F, rng)) for _ in range(n_const)]
    public_ids = [W.new(_rand_ext(F, rng)) for _ in range(n_public)]

    def new_public(v):
        wid = W.new(v)
        public_ids.append(wid)
        return wid

    npo = (NpoTables(F, rng, W, const_ids + public_ids, new_public, n_perms, n_recompose, recompose_coeff)
           if (n_perms or n_recompose) else None)
    sources = const_ids + public_ids + (npo.created if npo else [])
    vals, preps13, pool = gen_alu_ops(F, rng, W, sources, n_alu)
    ops = finalize_alu(F, W, vals, preps13)

    def send_table(ids, lanes):
        v = np.array([W.values[i] for i in ids], dtype=np.uint32).reshape(-1, D)
        m = np.array([W.reads[i] for i in ids], dtype=np.uint32)
        idx = np.array(ids, dtype=np.uint32) * D
        return (witness_send.trace_to_matrix(v, D, lanes, min_height),
                witness_send.preprocessed_matrix(m, idx, lanes, min_height))

    tc, pc = send_table(const_ids, 1)
    tp, pp = send_table(public_ids, public_lanes)
    ta, pa = alu.build_tables(ops, F, D, alu_lanes, horner_k, min_height)
    lh = lambda m: int(m.shape[0]).bit_length() - 1
    aw, apw = alu.widths(D, alu_lanes, horner_k)
    insts = [
        air.build_instance("const", witness_send.make_eval(D, 1), F.p, lh(tc), D, 2, 0, buses),
        air.build_instance("public", witness_send.make_eval(D, public_lanes), F.p, lh(tp), D * public_lanes, 2 * public_lanes, 0, buses),
        air.build_instance("alu", alu.make_eval(D, alu_lanes, horner_k, F.w), F.p, lh(ta), aw, apw, 0, buses),
    ]
    prep_mats, traces, pubs = [pc, pp, pa], [tc, tp, ta], [None, None, None]
    p2_ops = {}
    if npo:
        ei, ep, et, epub = npo.finish(W, buses, min_height, recompose_lanes)
        p2_ops = {len(insts): npo.ops}
        insts += ei
        prep_mats += ep
        traces += et
        pubs += epub
    shapes = [(s.name, t.shape[0], t.shape[1], 0 if pm is None else pm.shape[1]) for s, t, pm in zip(insts, traces, prep_mats)]
    alu_ops = {2: alu.AluTableOps(ops, D, alu_lanes, horner_k)} if (D == 4 and horner_k == 4) else {}
    return LayerWorkload(insts, prep_mats, traces, pubs, shapes, p2_ops, alu_ops)


def base_layer_fibonacci(F: Field, n: int = 1000, min_height: int = 256) -> LayerWorkload:
    """The base circuit of `recursive_fibonacci` (recursion/examples/recursive_fibonacci.rs:315-327), extension degree 1,
    `TablePacking::new(1, 1)`: one public input `expected_result`, constants 0 and 1, n - 1 ADD operations, and
    `connect(b, expected_result)` — the last ADD's output IS the public witness, so that row reads it (`out_is_creator`
    false -> multiplicity -1, circuit-prover/src/common.rs:268-275) instead of creating it. Tables [Const, Public, ALU];
    for n = 1000: ALU 1024 x 7 with 20 preprocessed columns, Const / Public 256 x 1 (air/shape_golden.rs:33-52)."""
    p, d = F.p, 1
    values, reads = [0, 0, 1], [0, 0, 0]      # w0 = expected_result (value filled below), w1 = F(0), w2 = F(1)
    a_id, b_id = 1, 2
    vals, prep13 = [], []
    for step in range(2, n + 1):
        out = (values[a_id] + values[b_id]) % p
        last = step == n
        out_id = 0 if last else len(values)
        if last:
            values[0] = out
        else:
            values.append(out)
            reads.append(0)
        reads[a_id] += 1
        reads[b_id] += 1
        if last:
            reads[0] += 1
        pr = [0] * 13
        pr[alu.MULT_A] = pr[alu.MULT_B] = p - 1
        pr[alu.SEL_ADD], pr[alu.A_READER] = 1, 1
        pr[alu.A_IDX], pr[alu.B_IDX], pr[alu.OUT_IDX] = a_id * d, b_id * d, out_id * d
        pr[alu.MULT_OUT] = p - 1 if last else -1
        vals.append([[values[a_id]], [values[b_id]], [0], [out]])
        prep13.append(pr)
        a_id, b_id = b_id, out_id
    arr = np.zeros((len(prep13), 13), dtype=np.uint32)
    for i, pr in enumerate(prep13):
        if pr[alu.MULT_OUT] == -1:
            pr[alu.MULT_OUT] = reads[pr[alu.OUT_IDX] // d] % p
        arr[i] = pr
    ops = alu.AluOps(np.array(vals, dtype=np.uint32).reshape(-1, 4, d), arr)

    def send_table(ids):
        v = np.array([[values[i]] for i in ids], dtype=np.uint32)
        m = np.array([reads[i] for i in ids], dtype=np.uint32)
        idx = np.array(ids, dtype=np.uint32) * d
        return witness_send.trace_to_matrix(v, d, 1, min_height), witness_send.preprocessed_matrix(m, idx, 1, min_height)

    buses = air.BusRegistry()
    tc, pc = send_table([1, 2])
    tp, pp = send_table([0])
    ta, pa = alu.build_tables(ops, F, d, 1, 2, min_height)
    lh = lambda m: int(m.shape[0]).bit_length() - 1
    aw, apw = alu.widths(d, 1, 2)
    insts = [air.build_instance("const", witness_send.make_eval(d, 1), p, lh(tc), d, 2, 0, buses),
             air.build_instance("public", witness_send.make_eval(d, 1), p, lh(tp), d, 2, 0, buses),
             air.build_instance("alu", alu.make_eval(d, 1, 2, None), p, lh(ta), aw, apw, 0, buses)]
    traces, preps = [tc, tp, ta], [pc, pp, pa]
    shapes = [(s.name, t.shape[0], t.shape[1], pm.shape[1]) for s, t, pm in zip(insts, traces, preps)]
    return LayerWorkload(insts, preps, traces, [None, None, None], shapes)
