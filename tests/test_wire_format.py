"""Proof wire format: flat proof blob <-> postcard bytes of `BatchStarkProof` (csrc/wire.cpp, host-only).
The blob comes from the oracle prover (same layout as the CUDA path's; the GPU parity tests pin blob equality)."""
import importlib

import numpy as np
import pytest

from common import SMALL_FRI, field_mod, make_oracle

lib = importlib.import_module("plonky3-recursion_b200.lib")
wl = importlib.import_module("plonky3-recursion_b200.workload")


def _layer(field_name, fri, **kw):
    F = field_mod.get_field(field_name)
    L = wl.synthetic_layer(F, 21, n_const=9, n_public=17, n_alu=120, n_perms=24, n_recompose=5, min_height=16, **kw)
    orc = make_oracle(field_name, fri)
    blob = orc.prove(L.insts, L.preps, L.traces, L.pubs)
    cap = orc.prep_commit(L.insts, L.preps)
    meta = dict(public_lanes=1, alu_lanes=3, npo_lanes=[("recompose", 1)], min_trace_height=16, horner_packed_steps=4,
                rows=(9, 17, 120), ext_degree=4,
                non_primitives=[(f"poseidon2_perm/{'koala' if F.field_id == 0 else 'baby'}_bear_d4_w16", 24, 1, [], 0),
                                ("recompose", 5, 1, [], 0)], prep_cap=cap)
    return F, L, blob, meta


def _varint(n):
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


@pytest.mark.parametrize("field_name", ["koala-bear", "baby-bear"])
@pytest.mark.parametrize("flags", [0, lib.WIRE_CANONICAL, lib.WIRE_BARE_ROOT, lib.WIRE_CANONICAL | lib.WIRE_BARE_ROOT])
def test_blob_bytes_blob_round_trip(field_name, flags):
    F, L, blob, meta = _layer(field_name, SMALL_FRI)
    data, proof_len = lib.serialize_proof(F, SMALL_FRI, L.insts, blob, meta, flags)
    back, off = lib.deserialize_proof(F, SMALL_FRI, data, flags)
    assert off == proof_len and np.array_equal(back, blob)
    # byte count: every field element is one u32 varint (1..5 bytes), structure bytes are few
    n_fe = blob.size
    assert n_fe < len(data) <= 5 * n_fe + 1024
    # the metadata that follows `proof` starts with TablePacking { public_lanes: 1, alu_lanes: 3, npo_lanes: [("recompose", 1)], .. }
    assert data[proof_len:proof_len + 3] == bytes([1, 3, 1])
    assert data[proof_len + 3:proof_len + 13] == _varint(len("recompose")) + b"recompose"
    # a truncated byte string and a flipped length byte are rejected, not mis-parsed
    with pytest.raises(lib.P3RError):
        lib.deserialize_proof(F, SMALL_FRI, data[: proof_len // 2], flags)
    with pytest.raises(lib.P3RError):
        lib.serialize_proof(F, SMALL_FRI, L.insts, blob[:-3], meta, flags)


def test_known_prefix_of_the_encoding():
    """First bytes, spelled out by hand: commitments.main = cap Vec of one digest -> 0x01, 8 field elements as u32 varints of
    the Montgomery words; then Option tag 0x01 for the permutation commitment."""
    F, L, blob, meta = _layer("koala-bear", SMALL_FRI)
    data, _ = lib.serialize_proof(F, SMALL_FRI, L.insts, blob, meta, 0)
    n = len(L.insts)
    main_cap = blob[5 + n: 5 + n + 8]
    want = b"\x01" + b"".join(_varint(int(w)) for w in main_cap) + b"\x01\x01"
    assert data[: len(want)] == want
    canon, _ = lib.serialize_proof(F, SMALL_FRI, L.insts, blob, meta, lib.WIRE_CANONICAL | lib.WIRE_BARE_ROOT)
    want = b"".join(_varint(int(w)) for w in F.from_monty(main_cap)) + b"\x01"
    assert canon[: len(want)] == want


def test_cap_height_and_no_lookup_shapes():
    fri = dict(SMALL_FRI, cap_height=2)
    F, L, blob, meta = _layer("koala-bear", fri)
    data, proof_len = lib.serialize_proof(F, fri, L.insts, blob, meta, 0)
    assert data[0] == 4                      # Vec<[F; 8]> with 2^cap_height entries
    back, _ = lib.deserialize_proof(F, fri, data, 0)
    assert np.array_equal(back, blob)
    # base circuit: D = 1, no NPO tables
    Lb = wl.base_layer_fibonacci(F, 40, min_height=16)
    orc = make_oracle("koala-bear", SMALL_FRI)
    blob_b = orc.prove(Lb.insts, Lb.preps, Lb.traces, Lb.pubs)
    meta_b = dict(public_lanes=1, alu_lanes=1, npo_lanes=[], min_trace_height=16, horner_packed_steps=2, rows=(2, 1, 39),
                  ext_degree=1, non_primitives=[], prep_cap=orc.prep_commit(Lb.insts, Lb.preps))
    data_b, pl = lib.serialize_proof(F, SMALL_FRI, Lb.insts, blob_b, meta_b, 0)
    back_b, _ = lib.deserialize_proof(F, SMALL_FRI, data_b, 0)
    assert np.array_equal(back_b, blob_b)
    # ext_degree 1 -> w_binomial None: ... rows(3) alu_variant ext_degree=1 Option tag 0 quintic 0 non_primitives len 0
    assert data_b[pl:pl + 5] == bytes([1, 1, 0, 16, 2]) and data_b[pl + 5: pl + 13] == bytes([2, 1, 39, 0, 1, 0, 0, 0])


def test_deserializer_survives_mutated_bytes():
    """Child proofs arrive from other workers: p3r_proof_deserialize must answer damaged input (flipped bytes, truncation,
    oversized varints, inserted bytes) with an error or a blob, never with a crash or an out-of-bounds read."""
    F, L, blob, meta = _layer("koala-bear", SMALL_FRI)
    rng = np.random.default_rng(0)
    rejected = 0
    for flags in range(4):
        data, _ = lib.serialize_proof(F, SMALL_FRI, L.insts, blob, meta, flags)
        for it in range(80):
            b = bytearray(data)
            mode = it % 4
            if mode == 0:
                for _ in range(int(rng.integers(1, 6))):
                    b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            elif mode == 1:
                b = b[: int(rng.integers(0, len(b)))]
            elif mode == 2:
                pos = int(rng.integers(0, min(len(b), 200)))
                b[pos:pos + 5] = b"\xff\xff\xff\xff\x7f"
            else:
                pos = int(rng.integers(0, len(b)))
                b[pos:pos] = bytes(rng.integers(0, 256, size=int(rng.integers(1, 9)), dtype=np.uint8))
            try:
                lib.deserialize_proof(F, SMALL_FRI, bytes(b), flags)
            except lib.P3RError:
                rejected += 1
    assert rejected > 200          # truncations and structural damage are refused; a flipped payload byte may still parse
