"""Committed regression pins of the round-2 configurations (tests/golden/vectors_r2.json, written by
tests/golden/make_golden_r2.py from the oracle): width-24 permutation and leaf hasher, uni-stark transcript head, the eight
LogUp conventions, the postcard wire bytes under the four serde flag settings. Reproduced here by the oracle through BOTH of
its routes (plain per-column routines and the CPU-arm strips), so a change of either shows up on CPU. These are not reference
outputs (parity unpinned at the byte level)."""
import hashlib
import importlib
import json
import os
import sys

import numpy as np
import pytest

from common import SMALL_FRI, field_mod, make_oracle, p2mod

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
gen = importlib.import_module("make_golden_r2")
wl = importlib.import_module("plonky3-recursion_b200.workload")
sym = importlib.import_module("plonky3-recursion_b200.symbolic")
lib = importlib.import_module("plonky3-recursion_b200.lib")

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vectors_r2.json")))


def _routes(orc):
    """Both prove() routes where the fast one exists."""
    yield "plain", False
    if orc.set_fast_paths(True):
        yield "strips", True


@pytest.mark.parametrize("field_name", gen.FIELDS)
def test_width24_permutation_and_leaf_hasher(field_name):
    F = field_mod.get_field(field_name)
    g = GOLD["fields"][field_name]
    assert GOLD["fri"] == SMALL_FRI
    orc = make_oracle(field_name, SMALL_FRI)
    p24 = p2mod.Poseidon2Params(F.field_id, 24)
    st = np.array(g["perm24"]["inputs"], dtype=np.uint32)
    assert orc.poseidon2_permute_w(p24, st).tolist() == g["perm24"]["outputs"]
    assert p24.permute(st).tolist() == g["perm24"]["outputs"]                 # the numpy restatement agrees with the pin
    L = wl.synthetic_layer(F, g["leaf_hasher_w24"]["seed"], **g["leaf_hasher_w24"]["layer"])
    try:
        orc.set_leaf_hasher(p24)
        for _, fast in _routes(orc):
            orc.set_fast_paths(fast)
            assert gen.sha(orc.prove(L.insts, L.preps, L.traces, L.pubs)) == g["leaf_hasher_w24"]["proof_sha256"]
    finally:
        orc.set_leaf_hasher(None)
        orc.set_fast_paths(True)


@pytest.mark.parametrize("field_name", gen.FIELDS)
def test_uni_stark_proof(field_name):
    F = field_mod.get_field(field_name)
    g = GOLD["fields"][field_name]["uni_stark"]
    orc = make_oracle(field_name, SMALL_FRI)
    inst, t, pubs = gen.wide_instance(F, g["width"], g["log_n"])
    try:
        orc.set_uni_stark(True)
        for _, fast in _routes(orc):
            orc.set_fast_paths(fast)
            assert gen.sha(orc.prove([inst], [None], [t], [pubs])) == g["proof_sha256"]
    finally:
        orc.set_uni_stark(False)
        orc.set_fast_paths(True)


@pytest.mark.parametrize("field_name", gen.FIELDS)
def test_conventions_and_wire_bytes(field_name):
    F = field_mod.get_field(field_name)
    g = GOLD["fields"][field_name]
    orc = make_oracle(field_name, SMALL_FRI)
    saved = dict(sym.LOGUP_CONVENTIONS)
    try:
        seen = set()
        for entry in g["conventions"]["proofs"]:
            cv = entry["convention"]
            sym.LOGUP_CONVENTIONS.update(cv)
            orc.set_conventions(**cv)
            L = wl.synthetic_layer(F, g["conventions"]["seed"], **g["conventions"]["layer"])
            for _, fast in _routes(orc):
                orc.set_fast_paths(fast)
                assert gen.sha(orc.prove(L.insts, L.preps, L.traces, L.pubs)) == entry["proof_sha256"], cv
            seen.add(entry["proof_sha256"])
        assert len(seen) == 8                                                   # eight settings, eight different proofs
    finally:
        sym.LOGUP_CONVENTIONS.update(saved)
        orc.set_conventions(**gen.CONVENTIONS[0])
        orc.set_fast_paths(True)
    L = wl.synthetic_layer(F, g["wire"]["seed"], **g["wire"]["layer"])
    proof = orc.prove(L.insts, L.preps, L.traces, L.pubs)
    assert proof.size == g["wire"]["proof_words"]
    meta = gen.wire_meta(F, orc.prep_commit(L.insts, L.preps))
    for entry in g["wire"]["bytes"]:
        data, proof_len = lib.serialize_proof(F, SMALL_FRI, L.insts, proof, meta, entry["flags"])
        assert (len(data), proof_len) == (entry["n_bytes"], entry["proof_field_bytes"])
        assert hashlib.sha256(data).hexdigest() == entry["sha256"]
        back, _ = lib.deserialize_proof(F, SMALL_FRI, data, entry["flags"])
        assert np.array_equal(back, proof)
