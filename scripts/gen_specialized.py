#!/usr/bin/env python
"""Generate csrc/specialized_gen.cuh: straight-line CUDA versions of the quotient kernel for fixed constraint programs.

The generic `k_quotient` interprets the constraint bytecode (~25 SASS instructions of dispatch and local-memory traffic per
op). For the table AIRs of the recursion layer the program is static per (AIR, field, packing), so this script lowers the
*same bytecode* to straight-line code (slots become registers) that nvcc compiles at build time. At prep time the library
hashes the uploaded program (FNV-1a over the instruction words) and uses the specialised kernel when hash and field match,
otherwise the interpreter. Parity of both paths is tested (tests/test_gpu_parity.py).
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fm = importlib.import_module("plonky3-recursion_b200.field")
air = importlib.import_module("plonky3-recursion_b200.air")
sym = importlib.import_module("plonky3-recursion_b200.symbolic")
alu = importlib.import_module("plonky3-recursion_b200.airs.alu")
p2air = importlib.import_module("plonky3-recursion_b200.airs.poseidon2")
ws = importlib.import_module("plonky3-recursion_b200.airs.witness_send")
p2mod = importlib.import_module("plonky3-recursion_b200.poseidon2_params")


def fnv1a(words: np.ndarray) -> int:
    h = 0xCBF29CE484222325
    for w in words.reshape(-1).tolist():
        for k in range(4):
            h ^= (w >> (8 * k)) & 0xFF
            h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def monty_insns(F, prog):
    ins = np.array(prog.insns, dtype=np.uint32).reshape(-1, 4).copy()
    is_const = ins[:, 0] == sym.OP_B_CONST
    ins[is_const, 2] = F.to_monty(ins[is_const, 2])
    return ins


N_GROUPS = 8   # warps per CTA; each evaluates one group of constraints for the CTA's 32 rows
COST = {}


def op_cost(op):
    """Rough thread-instruction cost per op (balances the groups)."""
    if op in (sym.OP_B_MUL,):
        return 4
    if op in (sym.OP_B_ADD, sym.OP_B_SUB, sym.OP_B_NEG):
        return 2
    if op == sym.OP_E_MUL:
        return 110
    if op in (sym.OP_E_MULB,):
        return 16
    if op in (sym.OP_E_ADD, sym.OP_E_SUB, sym.OP_E_NEG):
        return 8
    if op == sym.OP_E_PERM:
        return 6
    if op == sym.OP_ASSERT_B:
        return 24
    if op == sym.OP_ASSERT_E:
        return 120
    return 2


B_DEFS = {sym.OP_B_MAIN, sym.OP_B_PREP, sym.OP_B_PUB, sym.OP_B_SEL, sym.OP_B_CONST, sym.OP_B_ADD, sym.OP_B_SUB, sym.OP_B_MUL,
          sym.OP_B_NEG}
E_DEFS = {sym.OP_E_PERM, sym.OP_E_CHAL, sym.OP_E_PVAL, sym.OP_E_CONST, sym.OP_E_FROMB, sym.OP_E_ADD, sym.OP_E_SUB, sym.OP_E_MUL,
          sym.OP_E_NEG, sym.OP_E_MULB, sym.OP_E_ADDB, sym.OP_E_SUBB}


def to_ssa(ins):
    """Slot program -> SSA: list of (op, value id or constraint index, operand refs). Operand refs are value ids for slot
    operands and raw immediates otherwise. Returns (nodes, kind) with kind[v] in 'b' / 'e'."""
    cur_b, cur_e = {}, {}
    nodes = []
    for op, d, x, y in ins.tolist():
        def B(k):
            return cur_b[k]

        def E(k):
            return cur_e[k]
        if op in (sym.OP_B_MAIN, sym.OP_B_PREP):
            src = ("imm", x, y)
        elif op in (sym.OP_B_PUB, sym.OP_B_SEL, sym.OP_B_CONST, sym.OP_E_CHAL, sym.OP_E_PVAL, sym.OP_E_CONST):
            src = ("imm", x, 0)
        elif op == sym.OP_E_PERM:
            src = ("imm", x, y)
        elif op in (sym.OP_B_ADD, sym.OP_B_SUB, sym.OP_B_MUL):
            src = ("val", B(x), B(y))
        elif op == sym.OP_B_NEG:
            src = ("val", B(x))
        elif op == sym.OP_E_FROMB:
            src = ("val", B(x))
        elif op in (sym.OP_E_ADD, sym.OP_E_SUB, sym.OP_E_MUL):
            src = ("val", E(x), E(y))
        elif op == sym.OP_E_NEG:
            src = ("val", E(x))
        elif op in (sym.OP_E_MULB, sym.OP_E_ADDB, sym.OP_E_SUBB):
            src = ("val", E(x), B(y))
        elif op == sym.OP_ASSERT_B:
            src = ("val", B(x))
        elif op == sym.OP_ASSERT_E:
            src = ("val", E(x))
        elif op == sym.OP_OUT_B:
            src = ("val", B(x))
        else:
            raise ValueError(op)
        vid = len(nodes)
        nodes.append((op, d, src))
        if op in B_DEFS:
            cur_b[d] = vid
        elif op in E_DEFS:
            cur_e[d] = vid
    return nodes


def split_groups(nodes, n_groups):
    """Contiguous groups of assertions with balanced (dependency-closed) cost; returns per group the sorted node ids."""
    asserts = [i for i, (op, _, _) in enumerate(nodes) if op in (sym.OP_ASSERT_B, sym.OP_ASSERT_E)]

    def closure(roots, seen):
        stack, new = list(roots), []
        while stack:
            v = stack.pop()
            if v in seen:
                continue
            seen.add(v)
            new.append(v)
            src = nodes[v][2]
            if src[0] == "val":
                stack.extend(src[1:])
        return new
    # marginal cost of each assertion in program order (what it adds to a group that already holds the previous ones) is
    # order dependent; a simple two-pass scheme: total cost with full sharing, then cut at equal shares
    total_seen = set()
    marg = []
    for a in asserts:
        marg.append(sum(op_cost(nodes[v][0]) for v in closure([a], total_seen)))
    total = sum(marg)
    groups, cur, acc, target = [], [], 0, total / n_groups
    for a, m in zip(asserts, marg):
        cur.append(a)
        acc += m
        if acc >= target * (len(groups) + 1) and len(groups) < n_groups - 1:
            groups.append(cur)
            cur = []
    groups.append(cur)
    while len(groups) < n_groups:
        groups.append([])
    return [sorted(closure(g, set())) for g in groups]


def emit_node(o, nodes, v, row, sel=None):
    op, d, src = nodes[v]
    if op == sym.OP_OUT_B:
        return
    if op == sym.OP_B_SEL and sel is not None:
        o.append(f"        const uint32_t b{v} = {sel[src[1]]};")
        return
    bn = lambda k: f"b{k}"
    en = lambda k: f"e{k}"
    if op == sym.OP_B_MAIN:
        o.append(f"        const uint32_t b{v} = __ldg(a.main + (size_t){src[1]} * cs + {row[src[2]]});")
    elif op == sym.OP_B_PREP:
        o.append(f"        const uint32_t b{v} = __ldg(a.prep + (size_t){src[1]} * cs + {row[src[2]]});")
    elif op == sym.OP_B_PUB:
        o.append(f"        const uint32_t b{v} = __ldg(a.pub + {src[1]});")
    elif op == sym.OP_B_SEL:
        o.append(f"        const uint32_t b{v} = a.sel[{src[1]}u * NQ + s];")
    elif op == sym.OP_B_CONST:
        o.append(f"        const uint32_t b{v} = {src[1]}u;")
    elif op == sym.OP_B_ADD:
        o.append(f"        const uint32_t b{v} = fadd<F>({bn(src[1])}, {bn(src[2])});")
    elif op == sym.OP_B_SUB:
        o.append(f"        const uint32_t b{v} = fsub<F>({bn(src[1])}, {bn(src[2])});")
    elif op == sym.OP_B_MUL:
        o.append(f"        const uint32_t b{v} = fmul<F>({bn(src[1])}, {bn(src[2])});")
    elif op == sym.OP_B_NEG:
        o.append(f"        const uint32_t b{v} = fneg<F>({bn(src[1])});")
    elif op == sym.OP_E_PERM:
        o.append(f"        Ext4 e{v}; {{ const uint32_t* p = a.perm + (size_t){4 * src[1]} * cs + {row[src[2]]}; "
                 f"e{v} = Ext4{{{{__ldg(p), __ldg(p + cs), __ldg(p + 2 * cs), __ldg(p + 3 * cs)}}}}; }}")
    elif op == sym.OP_E_CHAL:
        o.append(f"        const Ext4 e{v} = a.chal[{src[1]}];")
    elif op == sym.OP_E_PVAL:
        o.append(f"        const Ext4 e{v} = a.pval[{src[1]}];")
    elif op == sym.OP_E_CONST:
        o.append(f"        const Ext4 e{v} = a.econst[{src[1]}];")
    elif op == sym.OP_E_FROMB:
        o.append(f"        const Ext4 e{v} = ext_lift<F>({bn(src[1])});")
    elif op == sym.OP_E_ADD:
        o.append(f"        const Ext4 e{v} = eadd<F>({en(src[1])}, {en(src[2])});")
    elif op == sym.OP_E_SUB:
        o.append(f"        const Ext4 e{v} = esub<F>({en(src[1])}, {en(src[2])});")
    elif op == sym.OP_E_MUL:
        o.append(f"        const Ext4 e{v} = emul<F>({en(src[1])}, {en(src[2])}, wnr);")
    elif op == sym.OP_E_NEG:
        o.append(f"        const Ext4 e{v} = eneg<F>({en(src[1])});")
    elif op == sym.OP_E_MULB:
        o.append(f"        const Ext4 e{v} = emul_base<F>({en(src[1])}, {bn(src[2])});")
    elif op == sym.OP_E_ADDB:
        o.append(f"        const Ext4 e{v} = eadd_base<F>({en(src[1])}, {bn(src[2])});")
    elif op == sym.OP_E_SUBB:
        o.append(f"        const Ext4 e{v} = esub_base<F>({en(src[1])}, {bn(src[2])});")
    elif op == sym.OP_ASSERT_B:
        o.append(f"        acc = eadd<F>(acc, emul_base<F>(a.alpha_pows[{d}], {bn(src[1])}));")
    elif op == sym.OP_ASSERT_E:
        o.append(f"        acc = eadd<F>(acc, emul<F>(a.alpha_pows[{d}], {en(src[1])}, wnr));")
    else:
        raise ValueError(op)


def emit_kernel(name: str, fname: str, ins: np.ndarray, nb: int, ne: int) -> str:
    """One CTA = 32 quotient-domain rows x N_GROUPS warps; warp g folds the constraints of group g (the quotient is linear in
    the constraints, sum_k alpha^(N-1-k) c_k), the partial sums meet in shared memory. Four times the threads of a
    one-thread-per-row kernel and a quarter of the live registers each: the kernel is latency-bound, not throughput-bound."""
    nodes = to_ssa(ins)
    groups = split_groups(nodes, N_GROUPS)
    o = []
    o.append(f"__global__ void __launch_bounds__({32 * N_GROUPS}, 4) {name}(QuotientArgs a) {{")
    o.append(f"    using F = {fname};")
    o.append("    const uint32_t lq = a.log_n + a.log_qc, NQ = 1u << lq, n = 1u << a.log_n;")
    o.append("    const uint32_t g = threadIdx.x >> 5, lane = threadIdx.x & 31u;")
    o.append("    const uint32_t s_raw = blockIdx.x * 32 + lane;")
    o.append("    const uint32_t s = s_raw < NQ ? s_raw : NQ - 1;   // out-of-range lanes recompute the last row and drop it")
    o.append("    const uint32_t i = bitrev32(s, lq);")
    o.append("    const uint32_t r0 = s, r1 = bitrev32((i + (1u << a.log_qc)) & (NQ - 1), lq);")
    o.append("    const size_t cs = (size_t)n << a.log_blowup;")
    o.append("    const uint32_t wnr = a.wnr;")
    o.append("    Ext4 acc = ext_zero();")
    row = ["r0", "r1"]
    o.append("    switch (g) {")
    for gi, vs in enumerate(groups):
        o.append(f"    case {gi}: {{")
        for v in vs:
            emit_node(o, nodes, v, row)
        o.append("    } break;")
    o.append("    default: break;")
    o.append("    }")
    o.append(f"    __shared__ Ext4 part[{N_GROUPS}][32];")
    o.append("    part[g][lane] = acc;")
    o.append("    __syncthreads();")
    o.append("    if (g == 0 && s_raw < NQ) {")
    o.append(f"        for (int k = 1; k < {N_GROUPS}; k++) acc = eadd<F>(acc, part[k][lane]);")
    o.append("        Ext4 q = emul_base<F>(acc, a.inv_van[i & ((1u << a.log_qc) - 1)]);")
    o.append("        uint32_t c = i & ((1u << a.log_qc) - 1), r = i >> a.log_qc;")
    o.append("        for (int k = 0; k < 4; k++) a.chunks[((size_t)c * 4 + k) * n + r] = q.c[k];")
    o.append("    }")
    o.append("}")
    sizes = [len(v) for v in groups]
    print(name, "group sizes (SSA nodes):", sizes, "of", len(nodes))
    return "\n".join(o)


LOGUP_GROUPS = 4   # warps per CTA of the generated LogUp kernels; each handles a contiguous group of lookups for 32 rows


def logup_hash(ins, lookups, interactions) -> int:
    """FNV-1a over the lookup-input program and the lookup / interaction structure (the library recomputes it at prep time)."""
    words = list(ins.reshape(-1).tolist())
    for _, first, n in lookups:
        words += [first, n]
    for mult_out, elem_first, n_elems in interactions:
        words += [mult_out, elem_first, n_elems]
    return fnv1a(np.array(words, dtype=np.uint64))


def emit_logup_kernel(name: str, fname: str, ins: np.ndarray, lookups, interactions) -> str:
    """LogUp permutation-trace rows (k_logup_rows) as straight-line code. A CTA = 32 trace rows x LOGUP_GROUPS warps; warp g
    evaluates the tuple elements its lookups need (dead-code-eliminated program), the denominators
    prefix + sum_k beta^k * field_k, ONE extension inversion per lookup (its interactions share it), the fraction column, and
    a partial row sum; the partial sums meet in shared memory."""
    nodes = to_ssa(ins)
    out_node = {}
    for v, (op, d, src) in enumerate(nodes):
        if op == sym.OP_OUT_B:
            out_node[d] = src[1]
    n_lk = len(lookups)
    G = min(LOGUP_GROUPS, max(1, n_lk))
    # contiguous groups of lookups, balanced by interaction count
    total = sum(n for _, _, n in lookups)
    groups, cur, acc = [], [], 0
    for c, (_, first, n) in enumerate(lookups):
        cur.append(c)
        acc += n
        if acc >= total * (len(groups) + 1) / G and len(groups) < G - 1:
            groups.append(cur)
            cur = []
    groups.append(cur)
    while len(groups) < G:
        groups.append([])

    def closure(roots):
        seen, stack = set(), list(roots)
        while stack:
            v = stack.pop()
            if v in seen:
                continue
            seen.add(v)
            src = nodes[v][2]
            if src[0] == "val":
                stack.extend(src[1:])
        return sorted(seen)

    o = []
    o.append(f"__global__ void __launch_bounds__({32 * LOGUP_GROUPS}, 4) {name}(LogupArgs a) {{")
    o.append(f"    using F = {fname};")
    o.append("    const uint32_t n = 1u << a.log_n;")
    o.append("    const uint32_t g = threadIdx.x >> 5, lane = threadIdx.x & 31u;")
    o.append("    const uint32_t r_raw = blockIdx.x * 32 + lane;")
    o.append("    const uint32_t r0 = r_raw < n ? r_raw : n - 1, r1 = (r0 + 1) & (n - 1);")
    o.append("    const bool live = r_raw < n;")
    o.append("    const size_t cs = n;")
    o.append("    const uint32_t wnr = a.wnr;")
    o.append("    Ext4 tot = ext_zero();")
    o.append("    Ext4 bp[8];")
    o.append("    for (int k = 0; k < 8; k++) bp[k] = a.beta_pows[k];")
    row = ["r0", "r1"]
    sel = ["(r0 == 0 ? F::R : 0u)", "(r0 == n - 1 ? F::R : 0u)", "(r0 != n - 1 ? F::R : 0u)"]
    o.append("    switch (g) {")
    for gi, lks in enumerate(groups):
        o.append(f"    case {gi}: {{")
        roots = []
        for c in lks:
            _, first, nint = lookups[c]
            for j in range(first, first + nint):
                mult_out, elem_first, n_elems = interactions[j]
                roots.append(out_node[mult_out])
                roots += [out_node[elem_first + k] for k in range(n_elems)]
        for v in closure(roots):
            emit_node(o, nodes, v, row, sel)
        for c in lks:
            _, first, nint = lookups[c]
            o.append(f"        {{   // lookup {c}")
            o.append(f"            const Ext4 prefix = a.chal[{2 * c}];")
            for q, j in enumerate(range(first, first + nint)):
                mult_out, elem_first, n_elems = interactions[j]
                o.append(f"            Ext4 den{q} = prefix;")
                for k in range(n_elems):
                    o.append(f"            den{q} = eadd<F>(den{q}, emul_base<F>(bp[{k}], b{out_node[elem_first + k]}));")
            # shared inversion: prefix products
            o.append("            Ext4 acc = den0;")
            for q in range(1, nint):
                o.append(f"            const Ext4 pre{q} = acc;")
                o.append(f"            acc = emul<F>(acc, den{q}, wnr);")
            o.append("            Ext4 inv = einv<F>(acc, wnr);")
            o.append("            Ext4 frac = ext_zero();")
            for q in range(nint - 1, 0, -1):
                mult_out = interactions[first + q][0]
                o.append(f"            frac = eadd<F>(frac, emul_base<F>(emul<F>(inv, pre{q}, wnr), b{out_node[mult_out]}));")
                o.append(f"            inv = emul<F>(inv, den{q}, wnr);")
            o.append(f"            frac = eadd<F>(frac, emul_base<F>(inv, b{out_node[interactions[first][0]]}));")
            o.append("            if (live) {")
            o.append(f"                for (int k = 0; k < 4; k++) a.perm[(size_t)({4 * (c + 1)} + k) * n + r0] = frac.c[k];")
            o.append("            }")
            o.append("            tot = eadd<F>(tot, frac);")
            o.append("        }")
        o.append("    } break;")
    o.append("    default: break;")
    o.append("    }")
    o.append(f"    __shared__ Ext4 part[{LOGUP_GROUPS}][32];")
    o.append("    part[g][lane] = tot;")
    o.append("    __syncthreads();")
    o.append("    if (g == 0 && live) {")
    o.append(f"        for (int k = 1; k < {LOGUP_GROUPS}; k++) tot = eadd<F>(tot, part[k][lane]);")
    o.append("        a.rowsum[r0] = tot;")
    o.append("    }")
    o.append("}")
    print(name, "lookups per group:", [len(x) for x in groups], "of", n_lk)
    return "\n".join(o)


def main(out_path):
    kernels, registry, lk_registry = [], [], []
    for fname, cname in (("koala-bear", "KoalaBear"), ("baby-bear", "BabyBear")):
        F = fm.get_field(fname)
        prm = p2mod.Poseidon2Params(F.field_id)
        buses = air.BusRegistry()
        aw, apw = alu.widths(4, 3, 4)
        specs = [
            ("alu_d4_l3_k4", air.build_instance("alu", alu.make_eval(4, 3, 4, F.w), F.p, 8, aw, apw, 0, buses)),
            ("poseidon2_d4_w16", air.build_instance("p2", p2air.make_eval(prm), F.p, 8, *p2air.widths(prm), 0, buses)),
            # Const / Public (1 lane) and Recompose: no local constraints, only the LogUp ones
            ("send_d4_l1", air.build_instance("send", ws.make_eval(4, 1), F.p, 8, 4, 2, 0, buses)),
            ("recompose_d4_l1", air.build_instance("recompose", ws.make_eval(4, 1, idx_first=True), F.p, 8, 4, 2, 0, buses)),
            ("recompose_coeff_d4_l1", air.build_instance("recompose/coeff", ws.make_eval(4, 1, idx_first=True, coeff_lookups=True),
                                                         F.p, 8, 4, ws.prep_lane_width(4, True), 0, buses)),
        ]
        for tag, inst in specs:
            ins = monty_insns(F, inst.constraints)
            h = fnv1a(ins)
            kname = f"k_quotient_spec_{tag}_{cname}"
            kernels.append(emit_kernel(kname, cname, ins, inst.constraints.n_base_slots, inst.constraints.n_ext_slots))
            registry.append((h, F.field_id, kname, ins.shape[0]))
            if inst.lookup_inputs is not None and inst.lookups:
                lins = monty_insns(F, inst.lookup_inputs)
                lh = logup_hash(lins, inst.lookups, inst.interactions)
                lname = f"k_logup_spec_{tag}_{cname}"
                kernels.append(emit_logup_kernel(lname, cname, lins, inst.lookups, inst.interactions))
                lk_registry.append((lh, F.field_id, lname, lins.shape[0]))
    with open(out_path, "w") as f:
        f.write("// GENERATED by scripts/gen_specialized.py — do not edit. Straight-line quotient kernels for fixed constraint programs.\n")
        f.write("#pragma once\n#include \"spec.h\"\nnamespace p3r {\n\n")
        f.write("\n\n".join(kernels))
        f.write("\n\nconstexpr unsigned SPEC_ROWS_PER_CTA = 32, SPEC_THREADS = %d;\nstatic const SpecEntry SPEC_QUOTIENT[] = {\n" % (32 * N_GROUPS))
        for h, fid, kname, n in registry:
            f.write(f"    {{0x{h:016x}ull, {fid}, {n}u, {kname}}},\n")
        f.write("};\n\nconstexpr unsigned SPEC_LOGUP_THREADS = %d;\nstatic const SpecLogupEntry SPEC_LOGUP[] = {\n" % (32 * LOGUP_GROUPS))
        for h, fid, kname, n in lk_registry:
            f.write(f"    {{0x{h:016x}ull, {fid}, {n}u, {kname}}},\n")
        f.write("};\n\n}  // namespace p3r\n")
    print("wrote", out_path, [(hex(h), fid, k, n) for h, fid, k, n in registry])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "plonky3-recursion_b200", "csrc", "specialized_gen.cuh"))
