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


def emit_kernel(name: str, fname: str, ins: np.ndarray, nb: int, ne: int) -> str:
    o = []
    o.append(f"__global__ void __launch_bounds__(128, 4) {name}(QuotientArgs a) {{")
    o.append(f"    using F = {fname};")
    o.append("    const uint32_t lq = a.log_n + a.log_qc, NQ = 1u << lq, n = 1u << a.log_n;")
    o.append("    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;")
    o.append("    if (s >= NQ) return;")
    o.append("    const uint32_t i = bitrev32(s, lq);")
    o.append("    const uint32_t r0 = s, r1 = bitrev32((i + (1u << a.log_qc)) & (NQ - 1), lq);")
    o.append("    const size_t cs = (size_t)n << a.log_blowup;")
    o.append("    const uint32_t wnr = a.wnr;")
    o.append(f"    uint32_t b[{max(nb, 1)}];")
    o.append(f"    Ext4 e[{max(ne, 1)}];")
    o.append("    Ext4 acc = ext_zero();")
    row = ["r0", "r1"]
    for op, d, x, y in ins.tolist():
        if op == sym.OP_B_MAIN:
            o.append(f"    b[{d}] = __ldg(a.main + (size_t){x} * cs + {row[y]});")
        elif op == sym.OP_B_PREP:
            o.append(f"    b[{d}] = __ldg(a.prep + (size_t){x} * cs + {row[y]});")
        elif op == sym.OP_B_PUB:
            o.append(f"    b[{d}] = __ldg(a.pub + {x});")
        elif op == sym.OP_B_SEL:
            o.append(f"    b[{d}] = a.sel[{x}u * NQ + s];")
        elif op == sym.OP_B_CONST:
            o.append(f"    b[{d}] = {x}u;")
        elif op == sym.OP_B_ADD:
            o.append(f"    b[{d}] = fadd<F>(b[{x}], b[{y}]);")
        elif op == sym.OP_B_SUB:
            o.append(f"    b[{d}] = fsub<F>(b[{x}], b[{y}]);")
        elif op == sym.OP_B_MUL:
            o.append(f"    b[{d}] = fmul<F>(b[{x}], b[{y}]);")
        elif op == sym.OP_B_NEG:
            o.append(f"    b[{d}] = fneg<F>(b[{x}]);")
        elif op == sym.OP_E_PERM:
            o.append(f"    {{ const uint32_t* p = a.perm + (size_t){4 * x} * cs + {row[y]}; e[{d}] = Ext4{{{{__ldg(p), __ldg(p + cs), __ldg(p + 2 * cs), __ldg(p + 3 * cs)}}}}; }}")
        elif op == sym.OP_E_CHAL:
            o.append(f"    e[{d}] = a.chal[{x}];")
        elif op == sym.OP_E_PVAL:
            o.append(f"    e[{d}] = a.pval[{x}];")
        elif op == sym.OP_E_CONST:
            o.append(f"    e[{d}] = a.econst[{x}];")
        elif op == sym.OP_E_FROMB:
            o.append(f"    e[{d}] = ext_lift<F>(b[{x}]);")
        elif op == sym.OP_E_ADD:
            o.append(f"    e[{d}] = eadd<F>(e[{x}], e[{y}]);")
        elif op == sym.OP_E_SUB:
            o.append(f"    e[{d}] = esub<F>(e[{x}], e[{y}]);")
        elif op == sym.OP_E_MUL:
            o.append(f"    e[{d}] = emul<F>(e[{x}], e[{y}], wnr);")
        elif op == sym.OP_E_NEG:
            o.append(f"    e[{d}] = eneg<F>(e[{x}]);")
        elif op == sym.OP_E_MULB:
            o.append(f"    e[{d}] = emul_base<F>(e[{x}], b[{y}]);")
        elif op == sym.OP_E_ADDB:
            o.append(f"    e[{d}] = eadd_base<F>(e[{x}], b[{y}]);")
        elif op == sym.OP_E_SUBB:
            o.append(f"    e[{d}] = esub_base<F>(e[{x}], b[{y}]);")
        elif op == sym.OP_ASSERT_B:
            o.append(f"    acc = eadd<F>(acc, emul_base<F>(a.alpha_pows[{d}], b[{x}]));")
        elif op == sym.OP_ASSERT_E:
            o.append(f"    acc = eadd<F>(acc, emul<F>(a.alpha_pows[{d}], e[{x}], wnr));")
        else:
            raise ValueError(op)
    o.append("    Ext4 q = emul_base<F>(acc, a.inv_van[i & ((1u << a.log_qc) - 1)]);")
    o.append("    uint32_t c = i & ((1u << a.log_qc) - 1), r = i >> a.log_qc;")
    o.append("    for (int k = 0; k < 4; k++) a.chunks[((size_t)c * 4 + k) * n + r] = q.c[k];")
    o.append("}")
    return "\n".join(o)


def main(out_path):
    kernels, registry = [], []
    for fname, cname in (("koala-bear", "KoalaBear"), ("baby-bear", "BabyBear")):
        F = fm.get_field(fname)
        prm = p2mod.Poseidon2Params(F.field_id)
        buses = air.BusRegistry()
        aw, apw = alu.widths(4, 3, 4)
        specs = [
            ("alu_d4_l3_k4", air.build_instance("alu", alu.make_eval(4, 3, 4, F.w), F.p, 8, aw, apw, 0, buses)),
            ("poseidon2_d4_w16", air.build_instance("p2", p2air.make_eval(prm), F.p, 8, *p2air.widths(prm), 0, buses)),
        ]
        for tag, inst in specs:
            ins = monty_insns(F, inst.constraints)
            h = fnv1a(ins)
            kname = f"k_quotient_spec_{tag}_{cname}"
            kernels.append(emit_kernel(kname, cname, ins, inst.constraints.n_base_slots, inst.constraints.n_ext_slots))
            registry.append((h, F.field_id, kname, ins.shape[0]))
    with open(out_path, "w") as f:
        f.write("// GENERATED by scripts/gen_specialized.py — do not edit. Straight-line quotient kernels for fixed constraint programs.\n")
        f.write("#pragma once\n#include \"spec.h\"\nnamespace p3r {\n\n")
        f.write("\n\n".join(kernels))
        f.write("\n\nstatic const SpecEntry SPEC_QUOTIENT[] = {\n")
        for h, fid, kname, n in registry:
            f.write(f"    {{0x{h:016x}ull, {fid}, {n}u, {kname}}},\n")
        f.write("};\n\n}  // namespace p3r\n")
    print("wrote", out_path, [(hex(h), fid, k, n) for h, fid, k, n in registry])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "plonky3-recursion_b200", "csrc", "specialized_gen.cuh"))
