"""ctypes mirror of include/p3r.h (data-format structs + marshalling helpers).

Only the struct layouts live here; the product library loader is in `lib.py`. The CPU oracle under oracle/ reads the same
structs (they are the wire format of the boundary), but it is loaded only by tests/ and bench.py.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .field import Field
from .symbolic import OP_B_CONST, Program

u32 = C.c_uint32
u32p = C.POINTER(C.c_uint32)


class FieldDesc(C.Structure):
    _fields_ = [("field_id", u32), ("p", u32), ("w", u32), ("generator", u32)]


class Poseidon2Consts(C.Structure):
    _fields_ = [("width", u32), ("sbox_degree", u32), ("rounds_f", u32), ("rounds_p", u32),
                ("external_rc", u32p), ("internal_rc", u32p), ("internal_diag", u32p)]


class FriParams(C.Structure):
    _fields_ = [("log_blowup", u32), ("log_final_poly_len", u32), ("max_log_arity", u32), ("num_queries", u32),
                ("commit_pow_bits", u32), ("query_pow_bits", u32), ("cap_height", u32)]


class MatrixU32(C.Structure):
    _fields_ = [("data", u32p), ("height", u32), ("width", u32)]


class Insn(C.Structure):
    _fields_ = [("op", u32), ("dst", u32), ("a", u32), ("b", u32)]


class ProgramC(C.Structure):
    _fields_ = [("insns", C.POINTER(Insn)), ("n_insns", u32), ("n_base_slots", u32), ("n_ext_slots", u32),
                ("ext_consts", u32p), ("n_ext_consts", u32), ("n_constraints", u32), ("n_outputs", u32)]


class Poseidon2OpsC(C.Structure):
    _fields_ = [("n_ops", u32), ("input_values", u32p), ("mmcs_bit", C.POINTER(C.c_uint8)), ("mmcs_index_sum", u32p)]


class AluOpsC(C.Structure):
    _fields_ = [("lanes", u32), ("d", u32), ("k_max", u32), ("n_slots", u32), ("slot_kind", u32p), ("slot_first", u32p),
                ("n_ops", u32), ("values", u32p)]


class TableOpsC(C.Structure):
    _fields_ = [("poseidon2", C.POINTER(Poseidon2OpsC)), ("alu", C.POINTER(AluOpsC))]


class InteractionC(C.Structure):
    _fields_ = [("mult_out", u32), ("elem_out_first", u32), ("n_elems", u32)]


class LookupC(C.Structure):
    _fields_ = [("bus", u32), ("first_interaction", u32), ("n_interactions", u32)]


class InstanceDesc(C.Structure):
    _fields_ = [("log_height", u32), ("main_width", u32), ("prep_width", u32), ("n_public", u32),
                ("log_quotient_chunks", u32), ("uses_next_row", u32),
                ("constraints", ProgramC), ("lookup_inputs", ProgramC),
                ("lookups", C.POINTER(LookupC)), ("n_lookups", u32),
                ("interactions", C.POINTER(InteractionC)), ("n_interactions", u32)]


class Poseidon2ChainOpsC(C.Structure):
    _fields_ = [("n_rows", u32), ("new_start", C.POINTER(C.c_uint8)), ("merkle_path", C.POINTER(C.c_uint8)),
                ("mmcs_bit", C.POINTER(C.c_uint8)), ("witness_mask", C.POINTER(C.c_uint8)), ("values", u32p)]


class ConventionsC(C.Structure):
    _fields_ = [("logup_negate", u32), ("logup_first_power", u32), ("logup_descending", u32)]


class NpoEntryC(C.Structure):
    _fields_ = [("op_type", C.c_char_p), ("rows", C.c_uint64), ("lanes", C.c_uint64), ("public_values", u32p),
                ("n_public_values", u32), ("air_variant", u32)]


class ProofMetaC(C.Structure):
    _fields_ = [("public_lanes", C.c_uint64), ("alu_lanes", C.c_uint64), ("npo_lane_ops", C.POINTER(C.c_char_p)),
                ("npo_lane_counts", C.POINTER(C.c_uint64)), ("n_npo_lanes", u32), ("min_trace_height", C.c_uint64),
                ("horner_packed_steps", C.c_uint64), ("rows", C.c_uint64 * 3), ("alu_variant", u32), ("ext_degree", C.c_uint64),
                ("alu_quintic_trinomial", u32), ("non_primitives", C.POINTER(NpoEntryC)), ("n_non_primitives", u32),
                ("prep_cap", u32p)]


def as_u32p(a: np.ndarray):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u32p)


class Marshal:
    """Keeps numpy buffers alive for the lifetime of the ctypes structs that point into them."""

    def __init__(self, field: Field):
        self.field = field
        self._keep = []

    def keep(self, a):
        self._keep.append(a)
        return a

    def u32(self, a) -> np.ndarray:
        return self.keep(np.ascontiguousarray(a, dtype=np.uint32))

    def matrix(self, canonical: np.ndarray | None) -> MatrixU32:
        """canonical: (height, width) uint32 canonical residues -> Montgomery row-major."""
        if canonical is None:
            return MatrixU32(None, 0, 0)
        m = self.u32(self.field.to_monty(canonical))
        assert m.ndim == 2
        return MatrixU32(as_u32p(m), m.shape[0], m.shape[1])

    def matrix_monty(self, monty: np.ndarray) -> MatrixU32:
        m = self.u32(monty)
        return MatrixU32(as_u32p(m), m.shape[0], m.shape[1])

    def program(self, prog: Program | None) -> ProgramC:
        if prog is None:
            return ProgramC(None, 0, 1, 1, None, 0, 0, 0)
        ins = np.array(prog.insns, dtype=np.uint32).reshape(-1, 4).copy()
        is_const = ins[:, 0] == OP_B_CONST
        ins[is_const, 2] = self.field.to_monty(ins[is_const, 2])
        ins = self.u32(ins)
        ec = self.u32(self.field.to_monty(prog.ext_consts.reshape(-1))) if prog.ext_consts.size else self.u32(np.zeros(4))
        return ProgramC(C.cast(as_u32p(ins), C.POINTER(Insn)), ins.shape[0], prog.n_base_slots, prog.n_ext_slots,
                        as_u32p(ec), prog.ext_consts.shape[0], prog.n_constraints, prog.n_outputs)

    def fri(self, d: dict) -> FriParams:
        return FriParams(d["log_blowup"], d["log_final_poly_len"], d["max_log_arity"], d["num_queries"],
                         d["commit_pow_bits"], d["query_pow_bits"], d["cap_height"])

    def field_desc(self) -> FieldDesc:
        f = self.field
        return FieldDesc(f.field_id, f.p, f.w, f.generator)

    def poseidon2(self, params) -> Poseidon2Consts:
        f = self.field
        e = self.u32(f.to_monty(params.external_rc))
        i = self.u32(f.to_monty(params.internal_rc))
        d = self.u32(f.to_monty(params.internal_diag))
        return Poseidon2Consts(params.width, params.sbox_degree, params.rounds_f, params.rounds_p,
                               as_u32p(e), as_u32p(i), as_u32p(d))

    def instances(self, insts) -> C.Array:
        """insts: list of air.AirInstance -> contiguous array of InstanceDesc."""
        arr = (InstanceDesc * len(insts))()
        for k, s in enumerate(insts):
            lk = (LookupC * max(1, len(s.lookups)))()
            for j, (bus, first, n) in enumerate(s.lookups):
                lk[j] = LookupC(bus, first, n)
            it = (InteractionC * max(1, len(s.interactions)))()
            for j, (mo, ef, ne) in enumerate(s.interactions):
                it[j] = InteractionC(mo, ef, ne)
            self.keep(lk)
            self.keep(it)
            arr[k] = InstanceDesc(s.log_height, s.main_width, s.prep_width, s.n_public, s.log_quotient_chunks,
                                  int(s.uses_next_row), self.program(s.constraints), self.program(s.lookup_inputs),
                                  lk, len(s.lookups), it, len(s.interactions))
        return self.keep(arr)

    def matrices(self, mats) -> C.Array:
        arr = (MatrixU32 * len(mats))()
        for k, m in enumerate(mats):
            arr[k] = self.matrix(m)
        return self.keep(arr)

    def poseidon2_ops(self, ops_by_instance: dict, n_inst: int):
        """{instance index: airs.poseidon2.Poseidon2Ops} -> array of n_inst pointers to p3r_poseidon2_ops (NULL = matrix)."""
        arr = (C.POINTER(Poseidon2OpsC) * n_inst)()
        for i, ops in ops_by_instance.items():
            iv = self.u32(self.field.to_monty(ops.input_values.reshape(-1)))
            bit = self.keep(np.ascontiguousarray(ops.mmcs_bit, dtype=np.uint8))
            sm = self.u32(self.field.to_monty(ops.mmcs_index_sum))
            st = Poseidon2OpsC(ops.n, as_u32p(iv), bit.ctypes.data_as(C.POINTER(C.c_uint8)), as_u32p(sm))
            self.keep(st)
            arr[i] = C.pointer(st)
        return self.keep(arr)

    def table_ops(self, p2_by_instance: dict, alu_by_instance: dict, n_inst: int, alloc=None):
        """{instance: Poseidon2Ops}, {instance: airs.alu.AluTableOps} -> array of n_inst p3r_table_ops (both NULL = matrix).
        alloc(shape, dtype) -> array places the operand arrays (pinned host memory for the e2e path); default numpy."""
        def put(a, dtype=np.uint32):
            a = np.ascontiguousarray(a, dtype=dtype)
            if alloc is None:
                return self.keep(a)
            buf = alloc(a.shape, dtype)
            buf[...] = a
            return self.keep(buf)

        arr = (TableOpsC * n_inst)()
        for i, ops in (p2_by_instance or {}).items():
            iv = put(self.field.to_monty(ops.input_values.reshape(-1)))
            bit = put(ops.mmcs_bit, np.uint8)
            sm = put(self.field.to_monty(ops.mmcs_index_sum))
            st = self.keep(Poseidon2OpsC(ops.n, as_u32p(iv), bit.ctypes.data_as(C.POINTER(C.c_uint8)), as_u32p(sm)))
            arr[i].poseidon2 = C.pointer(st)
        for i, t in (alu_by_instance or {}).items():
            kind, first = put(t.slot_kind), put(t.slot_first)
            vals = put(self.field.to_monty(t.values.reshape(-1)))
            st = self.keep(AluOpsC(t.lanes, t.d, t.k_max, kind.size, as_u32p(kind), as_u32p(first), t.values.shape[0], as_u32p(vals)))
            arr[i].alu = C.pointer(st)
        return self.keep(arr)

    def proof_meta(self, meta: dict) -> ProofMetaC:
        """dict with the BatchStarkProof metadata (include/p3r.h p3r_proof_meta): public_lanes, alu_lanes, npo_lanes
        [(op_type, lanes)], min_trace_height, horner_packed_steps, rows (const, public, alu), ext_degree, non_primitives
        [(op_type, rows, lanes, canonical public values, air_variant)], prep_cap (Montgomery words or None)."""
        lanes = meta.get("npo_lanes", [])
        ops = (C.c_char_p * max(1, len(lanes)))(*[o.encode() for o, _ in lanes])
        cnt = (C.c_uint64 * max(1, len(lanes)))(*[int(n) for _, n in lanes])
        nps = meta.get("non_primitives", [])
        ents = (NpoEntryC * max(1, len(nps)))()
        for k, (op, rows, ln, pv, variant) in enumerate(nps):
            a = self.u32(self.field.to_monty(np.asarray(pv, dtype=np.uint32))) if len(pv) else self.u32(np.zeros(1))
            ents[k] = NpoEntryC(op.encode(), int(rows), int(ln), as_u32p(a), len(pv), int(variant))
        cap = meta.get("prep_cap")
        capa = self.u32(cap) if cap is not None else None
        self.keep((ops, cnt, ents))
        return ProofMetaC(int(meta["public_lanes"]), int(meta["alu_lanes"]), ops, cnt, len(lanes), int(meta["min_trace_height"]),
                          int(meta["horner_packed_steps"]), (C.c_uint64 * 3)(*[int(x) for x in meta["rows"]]),
                          int(meta.get("alu_variant", 0)), int(meta["ext_degree"]), int(meta.get("alu_quintic_trinomial", 0)),
                          ents, len(nps), as_u32p(capa) if capa is not None else None)

    def public_values(self, pubs, insts=None) -> C.Array:
        """pubs[k]: canonical public values of instance k (None / empty when it has none). With `insts` the lengths are
        checked against n_public here; the library checks for NULL again (P3R_ERR_INVALID_ARG)."""
        arr = (u32p * len(pubs))()
        for k, pv in enumerate(pubs):
            if insts is not None:
                have = 0 if pv is None else len(pv)
                if have != insts[k].n_public:
                    raise ValueError(f"instance {k}: {have} public values given, the AIR declares {insts[k].n_public}")
            if pv is None or len(pv) == 0:
                arr[k] = None
            else:
                a = self.u32(self.field.to_monty(np.asarray(pv, dtype=np.uint32)))
                arr[k] = as_u32p(a)
        return self.keep(arr)
