"""ctypes binding of libp3r_b200.so (include/p3r.h) — the product path. Fails loudly when the CUDA library or a GPU is
missing; there is no CPU fallback and nothing here touches oracle/.

Host-side mirror of the reference interface for this path:
  `ProverData.from_airs_and_degrees`  <->  ProverData::from_airs_and_degrees (/root/reference recursion/src/recursion.rs:376)
  `BatchStarkProver.prove_all_tables` <->  BatchStarkProver::prove_all_tables (circuit-prover/src/batch_stark_prover.rs:1203-1222)
"""
from __future__ import annotations

import time
import ctypes as C
import os
import subprocess

import numpy as np

from . import abi
from .field import get_field
from .poseidon2_params import Poseidon2Params

_HERE = os.path.dirname(os.path.abspath(__file__))
# P3R_LIB selects another BUILD of the same library (A/B timing of compile-time variants, scripts/ab_time.py); never the oracle.
_LIB_PATH = os.environ.get("P3R_LIB") or os.path.join(_HERE, "libp3r_b200.so")

DEFAULT_FRI = dict(log_blowup=2, log_final_poly_len=5, max_log_arity=2, num_queries=54, commit_pow_bits=0,
                   query_pow_bits=15, cap_height=0)  # /root/reference recursion/examples/recursive_fibonacci.rs:71-132


class P3RError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"p3r error {code}: {msg}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile csrc/ for sm_100a into libp3r_b200.so (in-tree)."""
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f.endswith((".cu", ".cuh", ".h", ".cpp")) or f == "Makefile"]
    srcs.append(os.path.join(_HERE, "..", "include", "p3r.h"))
    stale = not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-s"])
    return _LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise P3RError(-1, f"{_LIB_PATH} not built: run __graft_entry__.build() (no CPU fallback exists)")
        lib = C.CDLL(_LIB_PATH)
        lib.p3r_last_error.restype = C.c_char_p
        lib.p3r_last_error.argtypes = [C.c_void_p]
        lib.p3r_build_info.restype = C.c_char_p
        lib.p3r_launch_count.restype = C.c_uint64
        lib.p3r_launch_count.argtypes = [C.c_void_p]
        lib.p3r_abi_version.restype = C.c_uint32
        lib.p3r_host_alloc.restype = C.c_void_p
        lib.p3r_host_alloc.argtypes = [C.c_size_t]
        lib.p3r_host_free.restype = None
        lib.p3r_host_free.argtypes = [C.c_void_p]
        lib.p3r_host_hasher_permute.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.p3r_host_hasher_free.restype = None
        lib.p3r_host_hasher_free.argtypes = [C.c_void_p]
        lib.p3r_set_wait_mode.restype = None
        lib.p3r_set_wait_mode.argtypes = [C.c_int]
        for name in ("p3r_ctx_destroy", "p3r_prep_free", "p3r_session_free", "p3r_traces_free"):
            getattr(lib, name).restype = None
            getattr(lib, name).argtypes = [C.c_void_p]
        _lib = lib
    return _lib


EXPORTS = ["p3r_abi_version", "p3r_build_info", "p3r_ctx_create", "p3r_ctx_destroy", "p3r_last_error", "p3r_prep_commit",
           "p3r_prep_free", "p3r_prove_begin", "p3r_commit_main", "p3r_commit_perm", "p3r_commit_quotient", "p3r_open",
           "p3r_fri_begin", "p3r_fri_commit", "p3r_fri_fold", "p3r_fri_final_poly", "p3r_fri_query", "p3r_session_free",
           "p3r_grind", "p3r_prove", "p3r_coset_lde", "p3r_mmcs_commit", "p3r_poseidon2_permute", "p3r_bench_commit",
           "p3r_last_phase_times", "p3r_launch_count", "p3r_traces_upload", "p3r_traces_free", "p3r_prove_resident",
           "p3r_host_alloc", "p3r_host_free", "p3r_timer_start", "p3r_timer_stop", "p3r_set_kernel_timing",
           "p3r_reset_kernel_stats", "p3r_kernel_stats", "p3r_set_specialization", "p3r_prove_ex", "p3r_traces_upload_ex",
           "p3r_traces_download", "p3r_kernel_perms", "p3r_bench_fri_round", "p3r_prove_ops",
           "p3r_traces_upload_ops", "p3r_set_wait_mode", "p3r_ctx_set_stream_priority", "p3r_host_hasher_create", "p3r_host_hasher_permute",
           "p3r_host_hasher_free", "p3r_traces_write_rows", "p3r_proof_serialize", "p3r_proof_deserialize",
           "p3r_wire_last_error", "p3r_ctx_set_conventions", "p3r_ctx_set_leaf_hasher", "p3r_poseidon2_permute_w", "p3r_ctx_set_uni_stark", "p3r_bench_commit_multi", "p3r_poseidon2_run_chains"]

WIRE_CANONICAL, WIRE_BARE_ROOT = 1, 2


def serialize_proof(field, fri: dict, insts, blob: np.ndarray, meta: dict, flags: int = 0):
    """Flat proof blob -> postcard bytes of the reference's `BatchStarkProof` (p3r_proof_serialize; host-only, no GPU needed).
    Returns (bytes, length of the leading `proof: BatchProof` field)."""
    lib = load()
    lib.p3r_wire_last_error.restype = C.c_char_p
    F = get_field(field) if isinstance(field, str) else field
    m = abi.Marshal(F)
    fd, fp, descs, mc = m.field_desc(), m.fri(fri), m.instances(insts), m.proof_meta(meta)
    blob = np.ascontiguousarray(blob, dtype=np.uint32)
    out = np.zeros(blob.size * 5 + 4096, dtype=np.uint8)
    n, pb = C.c_size_t(0), C.c_size_t(0)
    rc = lib.p3r_proof_serialize(C.byref(fd), C.byref(fp), len(insts), descs, abi.as_u32p(blob), C.c_size_t(blob.size), C.byref(mc),
                                 flags, out.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_size_t(out.size), C.byref(n), C.byref(pb))
    if rc != 0:
        raise P3RError(rc, lib.p3r_wire_last_error().decode())
    return out[: n.value].tobytes(), pb.value


def deserialize_proof(field, fri: dict, data: bytes, flags: int = 0):
    """postcard bytes -> (flat proof blob, offset at which the BatchStarkProof metadata starts) (p3r_proof_deserialize)."""
    lib = load()
    lib.p3r_wire_last_error.restype = C.c_char_p
    F = get_field(field) if isinstance(field, str) else field
    m = abi.Marshal(F)
    fd, fp = m.field_desc(), m.fri(fri)
    buf = np.frombuffer(data, dtype=np.uint8)
    out = np.zeros(len(data) + 64, dtype=np.uint32)
    n, pb = C.c_size_t(0), C.c_size_t(0)
    rc = lib.p3r_proof_deserialize(C.byref(fd), C.byref(fp), buf.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_size_t(buf.size), flags,
                                   abi.as_u32p(out), C.c_size_t(out.size), C.byref(n), C.byref(pb))
    if rc != 0:
        raise P3RError(rc, lib.p3r_wire_last_error().decode())
    return out[: n.value].copy(), pb.value


KERNEL_CLASSES = ["ntt_lde", "hash_rows", "compress", "logup", "quotient", "open", "reduced_openings", "fri_fold", "transpose",
                  "misc"]


class HostHasher:
    """p3r_host_hasher: the prover's host-side Poseidon2 permutation (no GPU needed). Canonical states in, canonical out."""

    def __init__(self, field="koala-bear", poseidon2: Poseidon2Params | None = None):
        self.lib = load()
        self.field = get_field(field) if isinstance(field, str) else field
        self._m = abi.Marshal(self.field)
        fd, pc = self._m.field_desc(), self._m.poseidon2(poseidon2 or Poseidon2Params(self.field.field_id))
        h = C.c_void_p()
        rc = self.lib.p3r_host_hasher_create(C.byref(fd), C.byref(pc), C.byref(h))
        if rc != 0:
            raise P3RError(rc, "p3r_host_hasher_create failed")
        self.h = h

    def permute(self, states_canonical) -> np.ndarray:
        s = np.ascontiguousarray(self.field.to_monty(np.asarray(states_canonical, dtype=np.uint32)).reshape(-1, 16))
        rc = self.lib.p3r_host_hasher_permute(self.h, abi.as_u32p(s), s.shape[0])
        if rc != 0:
            raise P3RError(rc, "p3r_host_hasher_permute failed")
        return self.field.from_monty(s)

    def close(self):
        if getattr(self, "h", None):
            self.lib.p3r_host_hasher_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """p3r_ctx: one prover context on one CUDA device (replaces building StarkConfig/MyPcs,
    recursion/examples/common/mod.rs:464-486)."""

    def __init__(self, field="koala-bear", fri: dict | None = None, device: int = 0, poseidon2: Poseidon2Params | None = None):
        self.lib = load()
        self.field = get_field(field)
        self.fri = dict(DEFAULT_FRI if fri is None else fri)
        self.p2 = poseidon2 or Poseidon2Params(self.field.field_id)
        self._m = abi.Marshal(self.field)
        fd, pc, fp = self._m.field_desc(), self._m.poseidon2(self.p2), self._m.fri(self.fri)
        h = C.c_void_p()
        rc = self.lib.p3r_ctx_create(device, C.byref(fd), C.byref(pc), C.byref(fp), C.byref(h))
        if rc != 0:
            raise P3RError(rc, self.lib.p3r_last_error(None).decode() or "p3r_ctx_create failed")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.p3r_ctx_destroy(self.h)   # drains the stream first: nothing reads the pinned blocks afterwards
            self.h = None
            for ptr in getattr(self, "_pinned", []):
                self.lib.p3r_host_free(ptr)
            self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise P3RError(rc, self.lib.p3r_last_error(self.h).decode())

    @property
    def cap_words(self):
        return 8 << self.fri["cap_height"]

    def launch_count(self) -> int:
        return int(self.lib.p3r_launch_count(self.h))

    # ---- isolated kernels -------------------------------------------------------------------------
    def poseidon2_permute(self, states_canonical: np.ndarray) -> np.ndarray:
        s = np.ascontiguousarray(self.field.to_monty(states_canonical).reshape(-1, 16))
        self._check(self.lib.p3r_poseidon2_permute(self.h, abi.as_u32p(s), s.shape[0]))
        return self.field.from_monty(s)

    def coset_lde(self, mat_canonical: np.ndarray, log_blowup: int) -> np.ndarray:
        m = abi.Marshal(self.field)
        mm = m.matrix(mat_canonical)
        out = np.zeros((mat_canonical.shape[0] << log_blowup, mat_canonical.shape[1]), dtype=np.uint32)
        self._check(self.lib.p3r_coset_lde(self.h, C.byref(mm), log_blowup, abi.as_u32p(out)))
        return self.field.from_monty(out)

    def mmcs_commit(self, mats_canonical) -> np.ndarray:
        m = abi.Marshal(self.field)
        arr = m.matrices(mats_canonical)
        cap = np.zeros(self.cap_words, dtype=np.uint32)
        self._check(self.lib.p3r_mmcs_commit(self.h, len(mats_canonical), arr, abi.as_u32p(cap)))
        return cap

    def grind(self, state_monty, pending_monty, bits) -> int:
        st = np.ascontiguousarray(state_monty, dtype=np.uint32)
        pe = np.ascontiguousarray(pending_monty, dtype=np.uint32)
        w = C.c_uint32(0)
        self._check(self.lib.p3r_grind(self.h, abi.as_u32p(st), abi.as_u32p(pe) if pe.size else None, pe.size, bits, C.byref(w)))
        return w.value

    def bench_commit(self, log_height: int, width: int, iters: int = 5, seed: int = 0xB200):
        t = (C.c_float * 3)()
        self._check(self.lib.p3r_bench_commit(self.h, log_height, width, iters, C.c_uint64(seed), t))
        return {"lde_ms": t[0], "merkle_ms": t[1]}

    def bench_commit_multi(self, log_heights, widths, iters: int = 3, seed: int = 0xB200):
        lh = (C.c_uint32 * len(log_heights))(*log_heights)
        wd = (C.c_uint32 * len(widths))(*widths)
        t = (C.c_float * 2)()
        self._check(self.lib.p3r_bench_commit_multi(self.h, len(log_heights), lh, wd, iters, C.c_uint64(seed), t))
        return {"lde_ms": t[0], "merkle_ms": t[1]}

    def bench_fri_round(self, log_len: int, log_arity: int, iters: int = 5, seed: int = 0xB200):
        t = (C.c_float * 2)()
        self._check(self.lib.p3r_bench_fri_round(self.h, log_len, log_arity, iters, C.c_uint64(seed), t))
        return {"fold_ms": t[0], "commit_ms": t[1]}

    def set_wait_mode(self, mode: str):
        """'spin' | 'yield' | 'block' | 'sleep' (process-wide, p3r_set_wait_mode)."""
        self.lib.p3r_set_wait_mode({"spin": 0, "yield": 1, "block": 2, "sleep": 3}[mode])

    def set_stream_priority(self, high: bool):
        """Schedule this context's kernels ahead of default-priority contexts (p3r_ctx_set_stream_priority)."""
        self._check(self.lib.p3r_ctx_set_stream_priority(self.h, 1 if high else 0))

    def poseidon2_run_chains(self, new_start, merkle_path, mmcs_bit, witness_mask, values_canonical):
        """Runner assist (p3r_poseidon2_run_chains): returns (inputs, outputs), canonical (n, 16) arrays."""
        u8 = lambda a: np.ascontiguousarray(a, dtype=np.uint8)
        ns, mp, mb, wm = u8(new_start), u8(merkle_path), u8(mmcs_bit), u8(witness_mask)
        v = np.ascontiguousarray(self.field.to_monty(np.asarray(values_canonical, dtype=np.uint32).reshape(-1, 16)))
        n = v.shape[0]
        p8 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint8))
        ops = abi.Poseidon2ChainOpsC(n, p8(ns), p8(mp), p8(mb), p8(wm), abi.as_u32p(v))
        ins, outs = np.zeros((n, 16), dtype=np.uint32), np.zeros((n, 16), dtype=np.uint32)
        self._check(self.lib.p3r_poseidon2_run_chains(self.h, C.byref(ops), abi.as_u32p(ins), abi.as_u32p(outs)))
        return self.field.from_monty(ins), self.field.from_monty(outs)

    def set_uni_stark(self, on: bool):
        """One-table proofs with p3-uni-stark's transcript head (p3r_ctx_set_uni_stark)."""
        self._check(self.lib.p3r_ctx_set_uni_stark(self.h, int(bool(on))))

    def set_leaf_hasher(self, params24: Poseidon2Params | None):
        """Width-24 leaf hashing (p3r_ctx_set_leaf_hasher): PaddingFreeSponge<Perm24, 24, 16, 8> for every MMCS leaf row;
        None = back to the width-16 sponge."""
        if params24 is None:
            self._check(self.lib.p3r_ctx_set_leaf_hasher(self.h, None))
            return
        pc = self._m.poseidon2(params24)
        self._check(self.lib.p3r_ctx_set_leaf_hasher(self.h, C.byref(pc)))

    def poseidon2_permute_w(self, params: Poseidon2Params, states_canonical: np.ndarray) -> np.ndarray:
        """Isolated permutation kernel of width params.width (16 or 24) with the given constants."""
        m = abi.Marshal(self.field)
        pc = m.poseidon2(params)
        s = np.ascontiguousarray(self.field.to_monty(states_canonical).reshape(-1, params.width))
        self._check(self.lib.p3r_poseidon2_permute_w(self.h, C.byref(pc), abi.as_u32p(s), s.shape[0]))
        return self.field.from_monty(s)

    def set_conventions(self, **conv):
        """[P3-EXT] protocol conventions (include/p3r.h p3r_conventions): logup_negate, logup_first_power, logup_descending.
        The instances must have been built under the same symbolic.LOGUP_CONVENTIONS."""
        c = abi.ConventionsC(int(conv.get("logup_negate", 0)), int(conv.get("logup_first_power", 0)),
                             int(conv.get("logup_descending", 0)))
        self._check(self.lib.p3r_ctx_set_conventions(self.h, C.byref(c)))

    def set_specialization(self, enable: bool):
        self._check(self.lib.p3r_set_specialization(self.h, int(enable)))

    def timer_start(self):
        self._check(self.lib.p3r_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        self._check(self.lib.p3r_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def set_kernel_timing(self, classes):
        mask = 0
        for c in classes:
            mask |= 1 << KERNEL_CLASSES.index(c)
        self._check(self.lib.p3r_set_kernel_timing(self.h, mask))

    def reset_kernel_stats(self):
        self._check(self.lib.p3r_reset_kernel_stats(self.h))

    def kernel_stats(self) -> dict:
        n = len(KERNEL_CLASSES)
        ms, la, by = (C.c_double * n)(), (C.c_uint64 * n)(), (C.c_uint64 * n)()
        cnt = C.c_uint32(0)
        self._check(self.lib.p3r_kernel_stats(self.h, None, ms, la, by, n, C.byref(cnt)))
        pe = (C.c_uint64 * n)()
        self._check(self.lib.p3r_kernel_perms(self.h, pe, n))
        return {KERNEL_CLASSES[k]: {"ms": float(ms[k]), "launches": int(la[k]), "bytes": int(by[k]), "perms": int(pe[k])}
                for k in range(n)}

    def pinned_empty(self, shape, dtype=np.uint32) -> np.ndarray:
        """numpy view of cudaHostAlloc'd memory owned by this context: freed in close(), so the array must not be used after
        the context is closed (TraceBatch / BatchStarkProver hold it and are bound to the context anyway)."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = self.lib.p3r_host_alloc(C.c_size_t(max(nbytes, 16)))
        if not ptr:
            raise P3RError(3, "p3r_host_alloc failed")
        buf = (C.c_uint8 * max(nbytes, 16)).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(ptr)
        return arr

    def last_phase_times(self) -> dict:
        names = C.c_char_p()
        ms = (C.c_float * 32)()
        n = C.c_uint32(0)
        self._check(self.lib.p3r_last_phase_times(self.h, C.byref(names), ms, 32, C.byref(n)))
        out, addr = {}, C.cast(names, C.c_void_p).value
        for i in range(min(n.value, 32)):          # NUL-terminated entries back to back: read each up to its own NUL
            name = C.string_at(addr)
            addr += len(name) + 1
            out[name.decode()] = float(ms[i])
        return out


class ProverData:
    """Device-resident preprocessed commitment + compiled AIR programs for one circuit shape
    (p3r_prep; ProverData::from_airs_and_degrees)."""

    def __init__(self, ctx: Context, insts, prep_mats, handle, cap, has_prep):
        self.ctx, self.insts, self.prep_mats, self.h = ctx, insts, prep_mats, handle
        self.preprocessed_commitment = cap if has_prep else None  # Montgomery words

    @classmethod
    def from_airs_and_degrees(cls, ctx: Context, insts, prep_mats):
        """insts: list[air.AirInstance]; prep_mats: list of canonical (h, w) uint32 arrays or None per instance."""
        m = abi.Marshal(ctx.field)
        descs = m.instances(insts)
        pm = m.matrices(prep_mats)
        cap = np.zeros(ctx.cap_words, dtype=np.uint32)
        has_prep = C.c_uint32(0)
        h = C.c_void_p()
        t0 = time.perf_counter()
        ctx._check(ctx.lib.p3r_prep_commit(ctx.h, len(insts), descs, pm, C.byref(h), abi.as_u32p(cap), C.byref(has_prep)))
        pd = cls(ctx, insts, prep_mats, h, cap, bool(has_prep.value))
        pd.commit_ms = (time.perf_counter() - t0) * 1e3   # the C-ABI call alone (returns after the tree root is on the host)
        return pd

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.p3r_prep_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TraceBatch:
    """Traces marshalled once (Montgomery, row-major) so repeated proofs do not pay the numpy conversion.
    pinned=True places the Montgomery matrices in cudaHostAlloc'd memory (the e2e path of bench.py)."""

    def __init__(self, ctx: Context, traces, pubs, pinned: bool = False, p2_ops: dict | None = None,
                 alu_ops: dict | None = None, insts=None):
        """p2_ops: {instance index: Poseidon2Ops}, alu_ops: {instance index: airs.alu.AluTableOps}; those instances' traces
        are generated on the device from the operation lists and their entry in `traces` may be None."""
        self.ctx = ctx
        self.m = abi.Marshal(ctx.field)
        self.tops = (self.m.table_ops(p2_ops, alu_ops, len(traces), ctx.pinned_empty if pinned else None)
                     if (p2_ops or alu_ops) else None)
        skip = set(p2_ops or {}) | set(alu_ops or {})
        traces = [None if k in skip else t for k, t in enumerate(traces)]
        # Montgomery row-major host matrices by instance (pinned=True: cudaHostAlloc'd); write_rows patches them in place
        arr = (abi.MatrixU32 * len(traces))()
        self._host = {}
        for k, t in enumerate(traces):
            if t is None:
                arr[k] = abi.MatrixU32(None, 0, 0)
                continue
            if pinned:
                buf = ctx.pinned_empty(t.shape)
                buf[...] = ctx.field.to_monty(t)
            else:
                buf = self.m.u32(ctx.field.to_monty(t))
            self._host[k] = buf
            arr[k] = abi.MatrixU32(abi.as_u32p(buf), t.shape[0], t.shape[1])
        self.tm = arr
        self.pv = self.m.public_values(pubs, insts)   # insts given: lengths checked against n_public
        self.h2d_bytes = int(sum(int(t.size) * 4 for t in traces if t is not None))
        if p2_ops:
            self.h2d_bytes += int(sum(o.n * (64 + 4 + 1) for o in p2_ops.values()))
        if alu_ops:
            self.h2d_bytes += int(sum(o.h2d_bytes for o in alu_ops.values()))
        self.resident = None

    def upload(self, prover_data):
        """Make the traces device-resident (p3r_traces_upload)."""
        h = C.c_void_p()
        self.ctx._check(self.ctx.lib.p3r_traces_upload_ops(self.ctx.h, prover_data.h, self.tm, self.tops, C.byref(h)))
        self.resident = h
        return self

    def write_rows(self, prover_data, inst: int, row0: int, rows_canonical: np.ndarray):
        """Overwrite rows [row0, row0 + n) of instance `inst`'s main trace: in the resident copy (p3r_traces_write_rows) when
        uploaded, and in the host (pinned) matrix when this batch carries one for that instance."""
        rows = np.ascontiguousarray(self.ctx.field.to_monty(np.asarray(rows_canonical, dtype=np.uint32)))
        if rows.ndim != 2 or rows.shape[1] != prover_data.insts[inst].main_width:
            raise ValueError("write_rows: rows must be (n, main_width)")
        if self.resident is not None:
            self.ctx._check(self.ctx.lib.p3r_traces_write_rows(self.ctx.h, prover_data.h, self.resident, inst, row0,
                                                               rows.shape[0], abi.as_u32p(rows)))
        host = self._host.get(inst)
        if host is not None:
            host[row0:row0 + rows.shape[0]] = rows

    def download(self, prover_data, inst: int) -> np.ndarray:
        """Main trace of instance `inst` as held on the device (canonical, row-major) — parity checks of the GPU table fill."""
        s = prover_data.insts[inst]
        out = np.zeros((1 << s.log_height, s.main_width), dtype=np.uint32)
        self.ctx._check(self.ctx.lib.p3r_traces_download(self.ctx.h, prover_data.h, self.resident, inst, abi.as_u32p(out)))
        return self.ctx.field.from_monty(out)

    def close(self):
        if self.resident is not None and self.ctx.h:
            self.ctx.lib.p3r_traces_free(self.resident)
        self.resident = None


class BatchStarkProver:
    """Mirror of BatchStarkProver<SC> (circuit-prover/src/batch_stark_prover.rs:685-697)."""

    def __init__(self, ctx: Context, pinned_output: bool = False):
        self.ctx = ctx
        self._pinned_output = pinned_output
        self._buf = ctx.pinned_empty((1 << 22,)) if pinned_output else np.zeros(1 << 22, dtype=np.uint32)
        self.last_proof_words = 0

    def prove_resident(self, traces: "TraceBatch", prover_data: ProverData, copy: bool = True):
        """Prove from device-resident traces (TraceBatch.upload)."""
        ctx = self.ctx
        n = self._call_growing(lambda n: ctx.lib.p3r_prove_resident(
            ctx.h, prover_data.h, traces.resident, traces.pv, abi.as_u32p(self._buf), C.c_size_t(self._buf.size), C.byref(n)))
        return self._buf[:n].copy() if copy else self._buf[:n]

    def _call_growing(self, call) -> int:
        """Run a prove entry point; on P3R_ERR_BUFFER (needed size in n) grow the proof buffer once and retry."""
        n = C.c_size_t(0)
        rc = call(n)
        if rc == 6 and n.value > self._buf.size:
            self._buf = (self.ctx.pinned_empty((n.value,)) if self._pinned_output else np.zeros(n.value, dtype=np.uint32))
            rc = call(n)
        self.ctx._check(rc)
        self.last_proof_words = n.value
        return n.value

    def prove_all_tables(self, traces, prover_data: ProverData, public_values=None, copy: bool = True) -> np.ndarray:
        """traces: list of canonical (h, w) uint32 matrices in instance order, or a TraceBatch.
        Returns the flat proof blob (DESIGN.md "Proof blob")."""
        ctx = self.ctx
        if not isinstance(traces, TraceBatch):
            pubs = public_values if public_values is not None else [None] * len(traces)
            traces = TraceBatch(ctx, traces, pubs, insts=prover_data.insts)
        n = self._call_growing(lambda n: ctx.lib.p3r_prove_ops(
            ctx.h, prover_data.h, traces.tm, traces.tops, traces.pv, abi.as_u32p(self._buf), C.c_size_t(self._buf.size), C.byref(n)))
        return self._buf[:n].copy() if copy else self._buf[:n]
