"""ctypes loader for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs — never by the product package.
"""
from __future__ import annotations

import ctypes as C
import importlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_pkg = importlib.import_module("plonky3-recursion_b200")
abi = importlib.import_module("plonky3-recursion_b200.abi")


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "p3r.h")
    inc = os.path.join(_HERE, "direct_airs.inc")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(inc)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


class Oracle:
    def __init__(self, field, p2params, fri: dict):
        self.lib = C.CDLL(build())
        self.lib.orc_last_error.restype = C.c_char_p
        self.field, self.p2, self.fri = field, p2params, fri
        self.m = abi.Marshal(field)
        fd, pc, fp = self.m.field_desc(), self.m.poseidon2(p2params), self.m.fri(fri)
        self._check(self.lib.orc_init(C.byref(fd), C.byref(pc), C.byref(fp)))

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.orc_last_error().decode())

    def set_conventions(self, **conv):
        """Same names and meaning as the library's p3r_conventions (process-wide in the oracle)."""
        c = abi.ConventionsC(int(conv.get("logup_negate", 0)), int(conv.get("logup_first_power", 0)),
                             int(conv.get("logup_descending", 0)))
        self._check(self.lib.orc_set_conventions(C.byref(c)))

    def set_threads(self, n: int = 0) -> int:
        """OpenMP threads of the oracle's parallel loops (0: query only). Returns the count in effect."""
        return int(self.lib.orc_set_threads(int(n)))

    def set_fast_paths(self, on: bool) -> bool:
        """prove() through the strip-wise AVX2 / batched-inverse routes of fast_paths.inc (default where the CPU has AVX2)
        or through the plain per-column routines. Returns whether the fast routes are in effect. Same proofs either way."""
        return bool(self.lib.orc_set_fast_paths(int(bool(on))))

    def set_uni_stark(self, on: bool):
        self._check(self.lib.orc_set_uni_stark(int(bool(on))))

    def set_leaf_hasher(self, params24=None):
        """Width-24 leaf sponge (rate 16) for every MMCS leaf row, or None for the width-16 sponge (process-wide)."""
        if params24 is None:
            self._check(self.lib.orc_set_leaf_hasher(None))
            return
        m = abi.Marshal(self.field)
        pc = m.poseidon2(params24)
        self._check(self.lib.orc_set_leaf_hasher(C.byref(pc)))

    def poseidon2_permute_w(self, params, states_canonical: np.ndarray) -> np.ndarray:
        m = abi.Marshal(self.field)
        pc = m.poseidon2(params)
        s = np.ascontiguousarray(self.field.to_monty(states_canonical).reshape(-1, params.width))
        self._check(self.lib.orc_poseidon2_permute_w(C.byref(pc), abi.as_u32p(s), s.shape[0]))
        return self.field.from_monty(s)

    def poseidon2_permute(self, states_canonical: np.ndarray) -> np.ndarray:
        s = np.ascontiguousarray(self.field.to_monty(states_canonical).reshape(-1, 16))
        self._check(self.lib.orc_poseidon2_permute(abi.as_u32p(s), s.shape[0]))
        return self.field.from_monty(s)

    def coset_lde(self, mat_canonical: np.ndarray, log_blowup: int) -> np.ndarray:
        m = abi.Marshal(self.field)
        mm = m.matrix(mat_canonical)
        out = np.zeros((mat_canonical.shape[0] << log_blowup, mat_canonical.shape[1]), dtype=np.uint32)
        self._check(self.lib.orc_coset_lde(C.byref(mm), log_blowup, abi.as_u32p(out)))
        return self.field.from_monty(out)

    def coset_lde_strips(self, mat_canonical: np.ndarray, log_ext: int, natural: bool = False) -> np.ndarray:
        """coset_lde through the eight-column strip route (fast_paths.inc); natural=True: rows in natural order (the
        quotient-domain flavour). Raises where the route does not exist (no AVX2)."""
        m = abi.Marshal(self.field)
        mm = m.matrix(mat_canonical)
        out = np.zeros((mat_canonical.shape[0] << log_ext, mat_canonical.shape[1]), dtype=np.uint32)
        self._check(self.lib.orc_coset_lde_strips(C.byref(mm), log_ext, int(natural), abi.as_u32p(out)))
        return self.field.from_monty(out)

    def mmcs_commit(self, mats_canonical) -> np.ndarray:
        m = abi.Marshal(self.field)
        arr = m.matrices(mats_canonical)
        cap = np.zeros(8 << self.fri["cap_height"], dtype=np.uint32)
        self._check(self.lib.orc_mmcs_commit(len(mats_canonical), arr, abi.as_u32p(cap)))
        return cap  # Montgomery words

    def mmcs_open_verify(self, mats_canonical, index: int):
        m = abi.Marshal(self.field)
        arr = m.matrices(mats_canonical)
        self._check(self.lib.orc_mmcs_open_verify(len(mats_canonical), arr, index))

    def grind(self, state_monty, pending_monty, bits) -> int:
        st = np.ascontiguousarray(state_monty, dtype=np.uint32)
        pe = np.ascontiguousarray(pending_monty, dtype=np.uint32)
        w = C.c_uint32(0)
        self._check(self.lib.orc_grind(abi.as_u32p(st), abi.as_u32p(pe) if pe.size else None, pe.size, bits, C.byref(w)))
        return w.value

    def challenger_script(self, ops, inputs_monty):
        ops = np.ascontiguousarray(ops, dtype=np.uint32).reshape(-1, 2)
        inp = np.ascontiguousarray(inputs_monty, dtype=np.uint32)
        out = np.zeros(ops.shape[0] + 1, dtype=np.uint32)
        n = C.c_uint32(0)
        self._check(self.lib.orc_challenger_script(abi.as_u32p(ops), ops.shape[0], abi.as_u32p(inp) if inp.size else None,
                                                   abi.as_u32p(out), C.byref(n)))
        return out[: n.value]

    def prep_commit(self, insts, prep_mats) -> np.ndarray:
        m = abi.Marshal(self.field)
        descs = m.instances(insts)
        pm = m.matrices(prep_mats)
        cap = np.zeros(8 << self.fri["cap_height"], dtype=np.uint32)
        self._check(self.lib.orc_prep_commit(len(insts), descs, pm, abi.as_u32p(cap)))
        return cap

    def prove(self, insts, prep_mats, traces, pubs, cap_words=1 << 24) -> np.ndarray:
        m = abi.Marshal(self.field)
        descs, pm, tm, pv = m.instances(insts), m.matrices(prep_mats), m.matrices(traces), m.public_values(pubs)
        out = np.zeros(cap_words, dtype=np.uint32)
        n = C.c_size_t(0)
        self._check(self.lib.orc_prove(len(insts), descs, pm, tm, pv, abi.as_u32p(out), C.c_size_t(cap_words), C.byref(n)))
        return out[: n.value].copy()

    def prepare(self, insts, prep_mats, traces, pubs, cap_words=1 << 24):
        """Marshal once (numpy -> Montgomery matrices, ctypes descriptors); the returned callable runs orc_prove alone, so a
        timing loop around it measures the prover and not the marshalling."""
        m = abi.Marshal(self.field)
        descs, pm, tm, pv = m.instances(insts), m.matrices(prep_mats), m.matrices(traces), m.public_values(pubs)
        out = np.zeros(cap_words, dtype=np.uint32)
        n = C.c_size_t(0)

        def run():
            _keep = m  # noqa: F841 - owns the buffers the descriptors point into
            self._check(self.lib.orc_prove(len(insts), descs, pm, tm, pv, abi.as_u32p(out), C.c_size_t(cap_words), C.byref(n)))
            return out[: n.value]
        return run

    def verify(self, insts, prep_cap_monty, pubs, proof: np.ndarray):
        m = abi.Marshal(self.field)
        descs, pv = m.instances(insts), m.public_values(pubs)
        proof = np.ascontiguousarray(proof, dtype=np.uint32)
        pc = abi.as_u32p(np.ascontiguousarray(prep_cap_monty, dtype=np.uint32)) if prep_cap_monty is not None else None
        self._check(self.lib.orc_verify(len(insts), descs, pc, pv, abi.as_u32p(proof), C.c_size_t(proof.size)))

    # ---- one (local, next) row pair: constraint values through the bytecode and through the hard-coded evaluators ----
    def _two_rows(self, local, nxt):
        return np.ascontiguousarray(self.field.to_monty(np.stack([np.asarray(local), np.asarray(nxt)]).astype(np.uint32)))

    def eval_air_rows(self, inst, local, nxt, prep_local, prep_next, sel, pub=None) -> np.ndarray:
        """AIR-only constraint values (canonical, one per constraint) of the instance's bytecode on this row pair."""
        import copy
        s = copy.copy(inst)
        s.constraints = inst.air_only_constraints
        s.lookups, s.interactions, s.lookup_inputs = [], [], None
        m = abi.Marshal(self.field)
        descs = m.instances([s])
        m2, p2 = self._two_rows(local, nxt), self._two_rows(prep_local, prep_next)
        pv = m.u32(self.field.to_monty(np.asarray(pub if pub is not None else [0], dtype=np.uint32)))
        sl = m.u32(self.field.to_monty(np.asarray(sel, dtype=np.uint32)))
        n_c = s.constraints.n_constraints
        out = np.zeros(4 * n_c, dtype=np.uint32)
        n = C.c_uint32(0)
        self._check(self.lib.orc_eval_air_rows(C.byref(descs[0]), abi.as_u32p(m2), abi.as_u32p(p2), abi.as_u32p(pv), abi.as_u32p(sl),
                                               abi.as_u32p(out), n_c, C.byref(n)))
        vals = self.field.from_monty(out).reshape(-1, 4)
        assert not vals[:, 1:].any()          # base-field constraints only
        return vals[:, 0]

    def alu_eval_direct(self, d, lanes, k_max, local, nxt, prep_local, prep_next) -> np.ndarray:
        m2, p2 = self._two_rows(local, nxt), self._two_rows(prep_local, prep_next)
        out = np.zeros(4096, dtype=np.uint32)
        n = C.c_uint32(0)
        self._check(self.lib.orc_alu_eval_direct(d, lanes, k_max, abi.as_u32p(m2), m2.shape[1], abi.as_u32p(p2), p2.shape[1],
                                                 abi.as_u32p(out), out.size, C.byref(n)))
        return self.field.from_monty(out[: n.value])

    def poseidon2_eval_direct(self, sbox_registers, local, nxt, prep_local, prep_next, is_transition) -> np.ndarray:
        m2, p2 = self._two_rows(local, nxt), self._two_rows(prep_local, prep_next)
        out = np.zeros(4096, dtype=np.uint32)
        n = C.c_uint32(0)
        it = int(self.field.to_monty(np.array([is_transition], dtype=np.uint32))[0])
        self._check(self.lib.orc_poseidon2_eval_direct(sbox_registers, abi.as_u32p(m2), m2.shape[1], abi.as_u32p(p2), p2.shape[1], it,
                                                       abi.as_u32p(out), out.size, C.byref(n)))
        return self.field.from_monty(out[: n.value])

    def check_constraints(self, inst, prep_mat, trace, pub):
        """AIR-only constraint check on the trace domain; returns (bad_row, bad_constraint) or None."""
        import copy
        s = copy.copy(inst)
        s.constraints = inst.air_only_constraints
        s.lookups, s.interactions, s.lookup_inputs = [], [], None
        if s.constraints is None:
            return None
        m = abi.Marshal(self.field)
        descs = m.instances([s])
        pm = m.matrix(prep_mat) if prep_mat is not None else abi.MatrixU32(None, 0, 0)
        tm = m.matrix(trace)
        pv = m.u32(self.field.to_monty(np.asarray(pub if pub is not None else [], dtype=np.uint32)))
        br, bc = C.c_int64(-1), C.c_int64(-1)
        rc = self.lib.orc_check_constraints(C.byref(descs[0]), C.byref(pm), C.byref(tm), abi.as_u32p(pv) if pv.size else None,
                                            C.byref(br), C.byref(bc))
        if rc != 0:
            if br.value >= 0:
                return (br.value, bc.value)
            raise RuntimeError(self.lib.orc_last_error().decode())
        return None
