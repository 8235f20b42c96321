"""AluAir: unified ADD / MUL / BOOL_CHECK / MUL_ADD / HORNER_ACC table with K-step packed Horner.

Restates /root/reference circuit-prover/src/air/alu_air.rs:
  layout doc :22-58, columns air/alu_columns.rs:8-46, `compute_schedule` :349-463, `trace_to_matrix` :497-608,
  `build_scheduled_preprocessed_trace` :610-681, `eval` :764-996, interactions :1000-1085;
per-op preprocessed values follow circuit-prover/src/common.rs:196-300 (13 columns per op).
Shape goldens (air/shape_golden.rs:47-61): D1 lane1 k2 -> (7, 20); D4 lane1 -> (28, 20); D4 lane2 -> (44, 33);
recursion layers use D4, 3 lanes, k=4 -> (80, 60).
"""
from __future__ import annotations

import numpy as np

PREP_LANE_WIDTH = 13
STEP_PREP_WIDTH = 6
# AluPrepLaneCols offsets
MULT_A, SEL_ADD, SEL_BOOL, SEL_MULADD, SEL_HORNER, A_IDX, B_IDX, C_IDX, OUT_IDX, MULT_B, MULT_OUT, A_READER, C_READER = range(13)

ADD, MUL, BOOL, MULADD, HORNER = range(5)


def num_horner_intermediates(k_max):
    return (k_max - 1) // 2


def horner_extra_prep_width(k):
    return (k - 1) + STEP_PREP_WIDTH * (k - 1)


def extra_prep_sel_k_idx(k):
    return k - 2


def extra_prep_a_idx_for_step(t, k_max):
    return (k_max - 1) + STEP_PREP_WIDTH * (t - 1)


def widths(d, lanes, k_max=2):
    main = lanes * 4 * d + (num_horner_intermediates(k_max) + 2 * (k_max - 1) + 1) * d
    prep = lanes * PREP_LANE_WIDTH + horner_extra_prep_width(k_max)
    return main, prep


def make_eval(d: int, lanes: int, k_max: int, w: int | None):
    """Returns eval(builder) mirroring `impl Air for AluAir` (binomial extension x^d = w, or base field when d == 1)."""
    lane_width = 4 * d

    def ext_mul(x, y):  # ext_mul_binomial, alu_air.rs:715-733
        acc = [None] * d
        for i in range(d):
            for j in range(d):
                term = x[i] * y[j]
                k = i + j
                if k >= d:
                    term = w * term
                    k -= d
                acc[k] = term if acc[k] is None else acc[k] + term
        return acc

    def eval_air(b):
        local = [b.main(c, 0) for c in range(b.main_width)]
        nxt = [b.main(c, 1) for c in range(b.main_width)]
        prep_local = [b.prep(c, 0) for c in range(b.prep_width)]
        prep_next = [b.prep(c, 1) for c in range(b.prep_width)]

        # ---- interactions first (eval_alu_interactions, :1000-1085) ----
        for lane in range(lanes):
            m, p = lane * lane_width, lane * PREP_LANE_WIDTH
            pl = prep_local[p:p + PREP_LANE_WIDTH]
            eff_a = pl[MULT_A] * pl[A_READER]
            eff_c = pl[MULT_A] * pl[C_READER]
            mults = [eff_a, pl[MULT_B], eff_c, pl[MULT_OUT]]
            idxs = [pl[A_IDX], pl[B_IDX], pl[C_IDX], pl[OUT_IDX]]
            for i in range(4):
                fields = [idxs[i]] + local[m + i * d:m + (i + 1) * d]
                b.push_interaction("WitnessChecks", fields, mults[i])
        extra_main, extra_prep = lanes * lane_width, lanes * PREP_LANE_WIDTH
        num_int = num_horner_intermediates(k_max)
        ac_base = extra_main + num_int * d
        for t in range(1, k_max):
            p = extra_prep + extra_prep_a_idx_for_step(t, k_max)
            step = prep_local[p:p + STEP_PREP_WIDTH]  # a_idx, c_idx, a_reader, c_reader, lookup_mult_a, lookup_mult_c
            off = ac_base + 2 * (t - 1) * d
            b.push_interaction("WitnessChecks", [step[0]] + local[off:off + d], step[4])
            b.push_interaction("WitnessChecks", [step[1]] + local[off + d:off + 2 * d], step[5])

        # ---- constraints (:764-996) ----
        for lane in range(lanes):
            m, p = lane * lane_width, lane * PREP_LANE_WIDTH
            a, bb_, c, out = (local[m + i * d:m + (i + 1) * d] for i in range(4))
            na, nb, nc, nout = (nxt[m + i * d:m + (i + 1) * d] for i in range(4))
            pc, pn = prep_local[p:p + PREP_LANE_WIDTH], prep_next[p:p + PREP_LANE_WIDTH]
            sel_add, sel_bool, sel_muladd, sel_horner = pc[SEL_ADD], pc[SEL_BOOL], pc[SEL_MULADD], pc[SEL_HORNER]
            active = 0 - pc[MULT_A]
            sel_mul = active - sel_bool - sel_muladd - sel_horner - sel_add
            for i in range(d):
                b.assert_zero(sel_add * (a[i] + bb_[i] - out[i]))
            ab = ext_mul(a, bb_)
            for i in range(d):
                b.assert_zero(sel_mul * (ab[i] - out[i]))
            b.assert_zero(sel_bool * a[0] * (a[0] - 1))
            for i in range(1, d):
                b.assert_zero(sel_bool * a[i])
            for i in range(d):
                b.assert_zero(sel_muladd * (ab[i] + c[i] - out[i]))
            next_sel_horner = pn[SEL_HORNER]
            out_next_b = ext_mul(out, nb)
            if lane == 0:
                next_int0 = nxt[extra_main:extra_main + d]
                any_cur = sum((prep_local[extra_prep + extra_prep_sel_k_idx(kk)] for kk in range(2, k_max + 1)), b.const(0))
                any_next = sum((prep_next[extra_prep + extra_prep_sel_k_idx(kk)] for kk in range(2, k_max + 1)), b.const(0))
                next_sel_k2 = prep_next[extra_prep + extra_prep_sel_k_idx(2)]
                sel_ge3_next = sum((prep_next[extra_prep + extra_prep_sel_k_idx(kk)] for kk in range(3, k_max + 1)), b.const(0))
                b_sq_base = ac_base + 2 * (k_max - 1) * d
                b_sq = local[b_sq_base:b_sq_base + d]
                b_sq_next = nxt[b_sq_base:b_sq_base + d]
                bsq_expr = ext_mul(bb_, bb_)
                for i in range(d):
                    b.assert_zero(any_cur * (b_sq[i] - bsq_expr[i]))
                out_b_sq = ext_mul(out, b_sq_next)
                c0_b_next = ext_mul(nc, nb)
                a0_b_next = ext_mul(na, nb)
                a1_next = nxt[ac_base:ac_base + d]
                c1_next = nxt[ac_base + d:ac_base + 2 * d]
                for i in range(d):
                    poly = out_b_sq[i] + c0_b_next[i] - a0_b_next[i] + c1_next[i] - a1_next[i]
                    b.assert_zero(next_sel_k2 * (poly - nout[i]))
                    b.assert_zero(sel_ge3_next * (poly - next_int0[i]))
                next_sel_single = next_sel_horner - any_next
                for i in range(d):
                    b.assert_zero(next_sel_single * (out_next_b[i] + nc[i] - na[i] - nout[i]))
                for kk in range(3, k_max + 1):
                    sel_kk = prep_local[extra_prep + extra_prep_sel_k_idx(kk)]
                    s, slot = 2, 0
                    while s < kk:
                        int_curr = local[extra_main + slot * d:extra_main + (slot + 1) * d]
                        off_s = ac_base + 2 * (s - 1) * d
                        a_s, c_s = local[off_s:off_s + d], local[off_s + d:off_s + 2 * d]
                        if s + 1 < kk:
                            off_sp1 = ac_base + 2 * s * d
                            a_sp1, c_sp1 = local[off_sp1:off_sp1 + d], local[off_sp1 + d:off_sp1 + 2 * d]
                            int_b_sq, c_s_b, a_s_b = ext_mul(int_curr, b_sq), ext_mul(c_s, bb_), ext_mul(a_s, bb_)
                            if s + 2 >= kk:
                                target = out
                            else:
                                target = local[extra_main + (slot + 1) * d:extra_main + (slot + 2) * d]
                                slot += 1
                            for i in range(d):
                                prod = int_b_sq[i] + c_s_b[i] - a_s_b[i] + c_sp1[i] - a_sp1[i]
                                b.assert_zero(sel_kk * (prod - target[i]))
                            s += 2
                        else:
                            int_b = ext_mul(int_curr, bb_)
                            for i in range(d):
                                b.assert_zero(sel_kk * (int_b[i] + c_s[i] - a_s[i] - out[i]))
                            s += 1
            else:
                for i in range(d):
                    b.assert_zero(next_sel_horner * (out_next_b[i] + nc[i] - na[i] - nout[i]))

    return eval_air


# ---------------------------------------------------------------------------------------------------
# Trace / preprocessed builders
# ---------------------------------------------------------------------------------------------------
class AluOps:
    """Logical ALU operations in circuit order.
    values: (n, 4, d) canonical [a, b, c, out]; prep13: (n, 13) canonical per-op preprocessed columns."""

    def __init__(self, values: np.ndarray, prep13: np.ndarray):
        self.values = np.asarray(values, dtype=np.uint32)
        self.prep13 = np.asarray(prep13, dtype=np.uint32)
        assert self.values.shape[0] == self.prep13.shape[0]


def compute_schedule(prep13: np.ndarray, lanes: int, pack_k: int):
    """alu_air.rs:349-463. Entries: ('op', i) | ('packed', first, k) | ('sep',). None when no HornerAcc op exists."""
    n = prep13.shape[0]
    if n == 0:
        return None
    is_h = prep13[:, SEL_HORNER] == 1
    if not is_h.any():
        return None
    chains, cur, non_chain = [], [], []
    for i in range(n):
        if is_h[i]:
            cur.append(i)
        else:
            if cur:
                chains.append(cur)
                cur = []
            non_chain.append(i)
    if cur:
        chains.append(cur)
    sched = []
    nc = [0]

    def fill_row():
        while len(sched) % lanes:
            if nc[0] < len(non_chain):
                sched.append(("op", non_chain[nc[0]]))
                nc[0] += 1
            else:
                sched.append(("sep",))

    sched.append(("sep",))
    fill_row()
    for ci, chain in enumerate(chains):
        if ci > 0:
            fill_row()
            sched.append(("sep",))
            fill_row()
        i = 0
        while i < len(chain):
            k_try = min(len(chain) - i, pack_k)
            best = 1
            for k in range(k_try, 1, -1):
                contiguous = all(chain[i + j] == chain[i] + j for j in range(1, k))
                if contiguous and len(set(int(prep13[chain[i + j], B_IDX]) for j in range(k))) == 1:
                    best = k
                    break
            if best >= 2:
                sched.append(("packed", chain[i], best))
                i += best
            else:
                sched.append(("op", chain[i]))
                i += 1
            fill_row()
    fill_row()
    while nc[0] < len(non_chain):
        sched.append(("op", non_chain[nc[0]]))
        nc[0] += 1
    fill_row()
    return sched


def _pad_height(rows: int, min_height: int) -> int:
    return max(min_height, 1 << max(rows - 1, 0).bit_length(), 1)


class AluTableOps:
    """What the device table fill needs (p3r_alu_ops): the schedule slots (static per circuit shape) and the operand values."""

    def __init__(self, ops: AluOps, d: int, lanes: int, k_max: int):
        sched = compute_schedule(ops.prep13, lanes, k_max)
        entries = sched if sched is not None else [("op", i) for i in range(ops.values.shape[0])]
        kind = np.zeros(len(entries), dtype=np.uint32)
        first = np.zeros(len(entries), dtype=np.uint32)
        for pos, ent in enumerate(entries):
            if ent[0] == "op":
                kind[pos], first[pos] = 1, ent[1]
            elif ent[0] == "packed":
                kind[pos], first[pos] = ent[2], ent[1]
        self.d, self.lanes, self.k_max = d, lanes, k_max
        self.slot_kind, self.slot_first = kind, first
        self.values = np.ascontiguousarray(ops.values, dtype=np.uint32)   # (n_ops, 4, d) canonical

    @property
    def h2d_bytes(self):
        return int(self.slot_kind.nbytes + self.slot_first.nbytes + self.values.nbytes)


def build_tables(ops: AluOps, field, d: int, lanes: int, k_max: int, min_height: int):
    """Returns (main matrix, preprocessed matrix), canonical uint32, padded with zero rows."""
    p = field.p
    main_w, prep_w = widths(d, lanes, k_max)
    lane_w = 4 * d
    n = ops.values.shape[0]
    sched = compute_schedule(ops.prep13, lanes, k_max)
    entries = sched if sched is not None else [("op", i) for i in range(n)]
    rows = -(-len(entries) // lanes) if entries else 0
    height = _pad_height(rows, min_height)
    main = np.zeros((height, main_w), dtype=np.uint32)
    prep = np.zeros((height, prep_w), dtype=np.uint32)
    extra_main, extra_prep = lanes * lane_w, lanes * PREP_LANE_WIDTH
    num_int = num_horner_intermediates(k_max)
    ac_base = extra_main + num_int * d
    b_sq_base = ac_base + 2 * (k_max - 1) * d

    def emul(x, y):
        return field.ext_mul(list(map(int, x)) + [0] * (4 - d), list(map(int, y)) + [0] * (4 - d))[:d] if d > 1 else [int(x[0]) * int(y[0]) % p]

    def eadd(x, y):
        return [(int(u) + int(v)) % p for u, v in zip(x, y)]

    def esub(x, y):
        return [(int(u) - int(v)) % p for u, v in zip(x, y)]

    prev_out = [0] * d
    for pos, ent in enumerate(entries):
        row, lane = divmod(pos, lanes)
        m, pp = lane * lane_w, lane * PREP_LANE_WIDTH
        if ent[0] == "op":
            i = ent[1]
            main[row, m:m + lane_w] = ops.values[i].reshape(-1)
            prep[row, pp:pp + PREP_LANE_WIDTH] = ops.prep13[i]
            if lane == 0:
                prev_out = list(ops.values[i, 3])
        elif ent[0] == "packed":
            first, k = ent[1], ent[2]
            last = first + k - 1
            main[row, m:m + 3 * d] = ops.values[first, :3].reshape(-1)
            main[row, m + 3 * d:m + 4 * d] = ops.values[last, 3]
            if lane == 0:
                bval = ops.values[first, 1]
                acc, step = prev_out, 0
                for s in range(num_int):
                    i0, i1 = first + step, first + step + 1
                    v0 = ops.values[i0]
                    if i1 < first + k:
                        v1 = ops.values[i1]
                        o0 = esub(eadd(emul(acc, bval), v0[2]), v0[0])
                        acc = esub(eadd(emul(o0, bval), v1[2]), v1[0])
                        step += 2
                    else:
                        acc = esub(eadd(emul(acc, bval), v0[2]), v0[0])
                        step += 1
                    main[row, extra_main + s * d:extra_main + (s + 1) * d] = acc
                for t in range(1, k):
                    off = ac_base + 2 * (t - 1) * d
                    main[row, off:off + d] = ops.values[first + t, 0]
                    main[row, off + d:off + 2 * d] = ops.values[first + t, 2]
                main[row, b_sq_base:b_sq_base + d] = emul(bval, bval)
                prev_out = list(main[row, 3 * d:4 * d])
                # preprocessed
                src0 = ops.prep13[first].copy()
                src_last = ops.prep13[last]
                src0[OUT_IDX] = src_last[OUT_IDX]
                src0[MULT_OUT] = src_last[MULT_OUT]
                src0[MULT_B] = int(src0[MULT_B]) * k % p
                prep[row, pp:pp + PREP_LANE_WIDTH] = src0
                mult_a_lane = int(src0[MULT_A])
                prep[row, extra_prep + extra_prep_sel_k_idx(k)] = 1
                for t in range(1, k):
                    st = ops.prep13[first + t]
                    q = extra_prep + extra_prep_a_idx_for_step(t, k_max)
                    prep[row, q + 0] = st[A_IDX]
                    prep[row, q + 1] = st[C_IDX]
                    prep[row, q + 2] = st[A_READER]
                    prep[row, q + 3] = st[C_READER]
                    prep[row, q + 4] = mult_a_lane * int(st[A_READER]) % p
                    prep[row, q + 5] = mult_a_lane * int(st[C_READER]) % p
        else:  # separator
            if lane == 0:
                prev_out = [0] * d
    return main, prep
