"""Poseidon2CircuitAir (D = 4, width 16, arity-2 Merkle shape): the Poseidon2 table of the recursion circuit.

Restates /root/reference poseidon2-circuit-air/src/air.rs:
  trace fill `generate_trace_rows` :280-520 (pass 1 inputs + mmcs_bit + index accumulator, pass 2 full round states),
  circuit-level constraints `eval` :832-1159 (non-compact branch :1024-1123), interactions :1798-1894,
  preprocessed padding :613-649; column wrapper poseidon-circuit-cols/src/cols.rs (perm | mmcs_bit | mmcs_index_sum),
  preprocessed row poseidon-circuit-cols/src/preprocessed.rs:30-200 (4 input limbs x [idx, in_ctl, normal_chain_sel,
  merkle_chain_sel], 2 output limbs x [idx, out_ctl], mmcs_index_sum_ctl_idx, mmcs_merkle_flag, new_start, merkle_path = 24).
The inner permutation AIR is p3-poseidon2-air's `Poseidon2Air` ([P3-EXT], recalled): columns
  inputs[16] | 4 x (sbox[16][R], post[16]) | P x (sbox[R], post_sbox) | 4 x (sbox[16][R], post[16])
with R = SBOX_REGISTERS (0 for KoalaBear degree 3, 1 for BabyBear degree 7) => 164 / 298 columns (+2 circuit columns:
166 / 300).
"""
from __future__ import annotations

import numpy as np

D = 4
WIDTH = 16
WIDTH_EXT = 4
RATE_EXT = 2
PREP_WIDTH = WIDTH_EXT * 4 + RATE_EXT * 2 + 4  # 24
# preprocessed tail offsets
MMCS_IDX, MMCS_FLAG, NEW_START, MERKLE_PATH = PREP_WIDTH - 4, PREP_WIDTH - 3, PREP_WIDTH - 2, PREP_WIDTH - 1


def sbox_registers(params) -> int:
    return {3: 0, 7: 1}[params.sbox_degree]


class Layout:
    def __init__(self, params):
        self.R = R = sbox_registers(params)
        self.half = params.rounds_f // 2
        self.rp = params.rounds_p
        c = WIDTH
        self.begin = []
        for _ in range(self.half):
            self.begin.append((c, c + WIDTH * R))  # (sbox base, post base)
            c += WIDTH * R + WIDTH
        self.partial = []
        for _ in range(self.rp):
            self.partial.append((c, c + R))        # (sbox base, post_sbox col)
            c += R + 1
        self.end = []
        for _ in range(self.half):
            self.end.append((c, c + WIDTH * R))
            c += WIDTH * R + WIDTH
        self.perm_cols = c
        self.mmcs_bit = c
        self.mmcs_index_sum = c + 1
        self.width = c + 2
        self.out_base = self.end[-1][1]            # ending_full_rounds[last].post


def widths(params):
    return Layout(params).width, PREP_WIDTH


def _external(state):
    """circ(2*M4, M4, M4, M4) on a list of 16 symbolic or numeric values (generic over + and small-constant *)."""
    out = [None] * 16
    for k in range(4):
        a, b, c, d = state[4 * k:4 * k + 4]
        t01, t23 = a + b, c + d
        t0123 = t01 + t23
        t01123, t01233 = t0123 + b, t0123 + d
        out[4 * k + 3] = t01233 + (a + a)
        out[4 * k + 1] = t01123 + (c + c)
        out[4 * k + 0] = t01123 + t01
        out[4 * k + 2] = t01233 + t23
    sums = [out[j] + out[4 + j] + out[8 + j] + out[12 + j] for j in range(4)]
    return [out[i] + sums[i % 4] for i in range(16)]


def make_eval(params):
    L = Layout(params)
    R, deg = L.R, params.sbox_degree
    erc = [int(x) for x in params.external_rc]
    irc = [int(x) for x in params.internal_rc]
    diag = [int(x) for x in params.internal_diag]

    def eval_air(b):
        local = [b.main(c, 0) for c in range(L.width)]
        nxt = [b.main(c, 1) for c in range(L.width)]
        pl = [b.prep(c, 0) for c in range(PREP_WIDTH)]
        pn = [b.prep(c, 1) for c in range(PREP_WIDTH)]
        local_out = local[L.out_base:L.out_base + WIDTH]
        next_in = nxt[0:WIDTH]
        next_bit = nxt[L.mmcs_bit]

        # ---- interactions (air.rs:1798-1894, arity-2 branch) ----
        not_merkle = 1 - pl[MERKLE_PATH]
        for limb in range(WIDTH_EXT):
            idx, in_ctl = pl[4 * limb + 0], pl[4 * limb + 1]
            fields = [idx] + local[limb * D:(limb + 1) * D]
            b.push_interaction("WitnessChecks", fields, 0 - in_ctl * not_merkle)
        ob = 4 * WIDTH_EXT
        for limb in range(RATE_EXT):
            idx, out_ctl = pl[ob + 2 * limb], pl[ob + 2 * limb + 1]
            fields = [idx] + local_out[limb * D:(limb + 1) * D]
            b.push_interaction("WitnessChecks", fields, out_ctl)
        mult = pl[MMCS_FLAG] * pn[NEW_START]
        b.push_interaction("WitnessChecks", [pl[MMCS_IDX], local[L.mmcs_index_sum], b.const(0), b.const(0), b.const(0)], 0 - mult)

        # ---- circuit-level constraints (air.rs:924-1123) ----
        b.assert_bool(local[L.mmcs_bit])
        tr = b.when_transition()
        for limb in range(WIDTH_EXT):
            gate = pn[4 * limb + 2]  # normal_chain_sel
            for d in range(D):
                tr.when(gate).assert_zero(next_in[limb * D + d] - local_out[limb * D + d])
        is_left = 1 - next_bit
        for i in range(RATE_EXT):
            gate_left = pn[4 * i + 3] * is_left
            for d in range(D):
                tr.when(gate_left).assert_zero(next_in[i * D + d] - local_out[i * D + d])
        for i in range(RATE_EXT):
            gate_right = pn[4 * i + 3] * next_bit
            for d in range(D):
                tr.when(gate_right).assert_zero(next_in[(RATE_EXT + i) * D + d] - local_out[i * D + d])
        not_next_new_start = 1 - pn[NEW_START]
        tr.when(not_next_new_start).when(pn[MERKLE_PATH]).assert_zero(
            nxt[L.mmcs_index_sum] - (local[L.mmcs_index_sum] * 2 + nxt[L.mmcs_bit]))

        # ---- inner Poseidon2Air::eval over the permutation columns ([P3-EXT] p3-poseidon2-air) ----
        def eval_sbox(regs, x):
            if R == 0:
                assert deg == 3
                return x * x * x
            committed_x3 = regs[0]
            b.assert_eq(committed_x3, x * x * x)
            return committed_x3 * committed_x3 * x

        state = _external(local[0:WIDTH])

        def full_round(state, sbox_base, post_base, rc):
            st = []
            for i in range(WIDTH):
                x = state[i] + rc[i]
                st.append(eval_sbox(local[sbox_base + i * R:sbox_base + (i + 1) * R], x))
            st = _external(st)
            for i in range(WIDTH):
                b.assert_eq(st[i], local[post_base + i])
            return local[post_base:post_base + WIDTH]

        for r in range(L.half):
            state = full_round(state, L.begin[r][0], L.begin[r][1], erc[16 * r:16 * r + 16])
        state = list(state)
        for r in range(L.rp):
            sb, post = L.partial[r]
            x = state[0] + irc[r]
            y = eval_sbox(local[sb:sb + R], x)
            b.assert_eq(y, local[post])
            state[0] = local[post]
            total = state[0]
            for i in range(1, WIDTH):
                total = total + state[i]
            state = [total + state[i] * diag[i] for i in range(WIDTH)]
        for r in range(L.half):
            state = full_round(state, L.end[r][0], L.end[r][1], erc[16 * (L.half + r):16 * (L.half + r) + 16])

    return eval_air


# ---------------------------------------------------------------------------------------------------
# Operations -> trace / preprocessed
# ---------------------------------------------------------------------------------------------------
class Poseidon2Ops:
    """Struct-of-arrays mirror of `Poseidon2CircuitRow` (circuit/src/ops/poseidon2_perm/trace.rs:94-124)."""

    def __init__(self, n):
        self.n = n
        self.new_start = np.zeros(n, dtype=bool)
        self.merkle_path = np.zeros(n, dtype=bool)
        self.mmcs_bit = np.zeros(n, dtype=bool)
        self.mmcs_index_sum = np.zeros(n, dtype=np.uint32)
        self.input_values = np.zeros((n, WIDTH), dtype=np.uint32)
        self.in_ctl = np.zeros((n, WIDTH_EXT), dtype=bool)
        self.input_indices = np.zeros((n, WIDTH_EXT), dtype=np.uint32)   # witness ids (unscaled)
        self.out_mult = np.zeros((n, RATE_EXT), dtype=np.uint32)        # creator multiplicity (0 = not exposed)
        self.output_indices = np.zeros((n, RATE_EXT), dtype=np.uint32)
        self.mmcs_index_sum_idx = np.zeros(n, dtype=np.uint32)
        self.mmcs_ctl_enabled = np.zeros(n, dtype=bool)


def padded_height(n_ops: int, min_height: int) -> int:
    return max(1 << max(n_ops - 1, 0).bit_length(), min_height, 1)


def permutation_states(params, inputs: np.ndarray):
    """All committed round values for `inputs` (n,16 canonical). Returns dict of arrays following
    p3_poseidon2_air::generate_trace_rows_for_perm."""
    p = np.uint64(params.field.p)
    L = Layout(params)
    n = inputs.shape[0]
    out = np.zeros((n, L.perm_cols), dtype=np.uint32)
    out[:, :WIDTH] = inputs
    erc = params.external_rc.astype(np.uint64).reshape(params.rounds_f, 16)
    diag = params.internal_diag.astype(np.uint64)
    s = params._external(inputs.astype(np.uint64) % p)

    def sbox(x):
        x2 = x * x % p
        x3 = x2 * x % p
        if params.sbox_degree == 3:
            return x3, x3
        return x3, (x3 * x3 % p) * x % p

    def full(s, bases, rc):
        x = (s + rc) % p
        x3, y = sbox(x)
        if L.R:
            out[:, bases[0]:bases[0] + WIDTH] = x3
        s = params._external(y)
        out[:, bases[1]:bases[1] + WIDTH] = s
        return s

    for r in range(L.half):
        s = full(s, L.begin[r], erc[r])
    for r in range(L.rp):
        x0 = (s[:, 0] + np.uint64(params.internal_rc[r])) % p
        x3, y = sbox(x0)
        if L.R:
            out[:, L.partial[r][0]] = x3
        out[:, L.partial[r][1]] = y
        s[:, 0] = y
        tot = s.sum(axis=1) % p
        s = (tot[:, None] + diag * s % p) % p
    for r in range(L.half):
        s = full(s, L.end[r], erc[L.half + r])
    return out


def build_tables(params, ops: Poseidon2Ops, min_height: int):
    """Main trace (Poseidon2CircuitAir::generate_trace_rows) and preprocessed trace, zero/`new_start` padded."""
    F = params.field
    L = Layout(params)
    n = ops.n
    H = padded_height(n, min_height)
    inputs = np.zeros((H, WIDTH), dtype=np.uint32)
    inputs[:n] = ops.input_values
    main = np.zeros((H, L.width), dtype=np.uint32)
    main[:, :L.perm_cols] = permutation_states(params, inputs)
    main[:n, L.mmcs_bit] = ops.mmcs_bit
    acc = np.zeros(H, dtype=np.uint64)
    prev = 0
    for r in range(n):  # pass 1 (sequential accumulator), air.rs:372-435
        if r > 0 and ops.merkle_path[r] and not ops.new_start[r]:
            prev = (prev * 2 + int(ops.mmcs_bit[r])) % F.p
        else:
            prev = int(ops.mmcs_index_sum[r])
        acc[r] = prev
    main[:, L.mmcs_index_sum] = acc.astype(np.uint32)
    prep = np.zeros((H, PREP_WIDTH), dtype=np.uint32)
    ns, mp = ops.new_start, ops.merkle_path
    for limb in range(WIDTH_EXT):
        ctl = ops.in_ctl[:, limb]
        prep[:n, 4 * limb + 0] = ops.input_indices[:, limb] * D
        prep[:n, 4 * limb + 1] = ctl
        prep[:n, 4 * limb + 2] = (~ns) & (~mp) & (~ctl)
        prep[:n, 4 * limb + 3] = (~ns) & mp & (~ctl)
    ob = 4 * WIDTH_EXT
    for limb in range(RATE_EXT):
        prep[:n, ob + 2 * limb] = ops.output_indices[:, limb] * D
        prep[:n, ob + 2 * limb + 1] = ops.out_mult[:, limb]
    prep[:n, MMCS_IDX] = ops.mmcs_index_sum_idx * D
    prep[:n, MMCS_FLAG] = ops.mmcs_ctl_enabled & mp
    prep[:n, NEW_START] = ns
    prep[:n, MERKLE_PATH] = mp
    if H > n:
        prep[n, NEW_START] = 1
    return main, prep
