#!/usr/bin/env python
"""Generate tests/golden/vectors_r2.json: regression pins (made by the CPU oracle, NOT reference outputs — see make_golden.py)
for the configurations added in round 2: the width-24 permutation and leaf hasher, the uni-stark transcript head, the eight
LogUp conventions, and the postcard wire bytes under the four serde flag settings.

Run from the repo root:  python tests/golden/make_golden_r2.py
"""
import hashlib
import importlib
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

from common import SMALL_FRI, air_mod, field_mod, make_oracle, p2mod  # noqa: E402

wl = importlib.import_module("plonky3-recursion_b200.workload")
sym = importlib.import_module("plonky3-recursion_b200.symbolic")
lib = importlib.import_module("plonky3-recursion_b200.lib")
wide = importlib.import_module("plonky3-recursion_b200.airs.wide")

FIELDS = ["koala-bear", "baby-bear"]
CONVENTIONS = [dict(logup_negate=n, logup_first_power=f, logup_descending=d) for n, f, d in itertools.product((0, 1), repeat=3)]
LAYER = dict(n_const=6, n_public=10, n_alu=60, n_perms=14, n_recompose=3, min_height=16)


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def wide_instance(F, width, log_n):
    t, pubs = wide.trace(F.p, width, log_n)
    inst = air_mod.build_instance("wide", wide.make_eval(width), F.p, log_n, width, 0, wide.N_PUBLIC, air_mod.BusRegistry())
    return inst, t, pubs


def wire_meta(F, cap):
    """BatchStarkProof metadata of the LAYER workload (TablePacking, rows, non-primitive table list)."""
    return dict(public_lanes=1, alu_lanes=3, npo_lanes=[("recompose", 1)], min_trace_height=16, horner_packed_steps=4,
                rows=(LAYER["n_const"], LAYER["n_public"], LAYER["n_alu"]), ext_degree=4,
                non_primitives=[(f"poseidon2_perm/{'koala' if F.field_id == 0 else 'baby'}_bear_d4_w16", LAYER["n_perms"], 1, [], 0),
                                ("recompose", LAYER["n_recompose"], 1, [], 0)], prep_cap=cap)


def vectors(field_name):
    F = field_mod.get_field(field_name)
    orc = make_oracle(field_name, SMALL_FRI)
    out = {}
    # width-24 permutation (canonical in, canonical out) and the width-24 leaf hasher's commitment / proof
    p24 = p2mod.Poseidon2Params(F.field_id, 24)
    st = np.array([[0] * 24, list(range(24)), [(i * 0x9E3779B1) % F.p for i in range(1, 25)]], dtype=np.uint32)
    out["perm24"] = {"inputs": st.tolist(), "outputs": orc.poseidon2_permute_w(p24, st).tolist()}
    L = wl.synthetic_layer(F, 4, **LAYER)
    try:
        orc.set_leaf_hasher(p24)
        out["leaf_hasher_w24"] = {"seed": 4, "layer": LAYER, "proof_sha256": sha(orc.prove(L.insts, L.preps, L.traces, L.pubs))}
    finally:
        orc.set_leaf_hasher(None)
    # uni-stark transcript head on a single wide table
    inst, t, pubs = wide_instance(F, 41, 5)
    try:
        orc.set_uni_stark(True)
        out["uni_stark"] = {"width": 41, "log_n": 5, "proof_sha256": sha(orc.prove([inst], [None], [t], [pubs]))}
    finally:
        orc.set_uni_stark(False)
    # the eight LogUp conventions (builder and prover switched together)
    saved = dict(sym.LOGUP_CONVENTIONS)
    conv = []
    try:
        for cv in CONVENTIONS:
            sym.LOGUP_CONVENTIONS.update(cv)
            orc.set_conventions(**cv)
            Lc = wl.synthetic_layer(F, 9, **LAYER)
            conv.append({"convention": cv, "proof_sha256": sha(orc.prove(Lc.insts, Lc.preps, Lc.traces, Lc.pubs))})
    finally:
        sym.LOGUP_CONVENTIONS.update(saved)
        orc.set_conventions(**CONVENTIONS[0])
    out["conventions"] = {"seed": 9, "layer": LAYER, "proofs": conv}
    # wire bytes of the default-convention proof under the four serde flag settings
    Lw = wl.synthetic_layer(F, 9, **LAYER)
    proof = orc.prove(Lw.insts, Lw.preps, Lw.traces, Lw.pubs)
    cap = orc.prep_commit(Lw.insts, Lw.preps)
    meta = wire_meta(F, cap)
    wire = []
    for flags in range(4):
        b, proof_len = lib.serialize_proof(F, SMALL_FRI, Lw.insts, proof, meta, flags)
        wire.append({"flags": flags, "n_bytes": len(b), "proof_field_bytes": int(proof_len),
                     "sha256": hashlib.sha256(bytes(b)).hexdigest()})
    out["wire"] = {"seed": 9, "layer": LAYER, "proof_words": int(proof.size), "bytes": wire}
    return out


def main():
    data = {"fri": SMALL_FRI, "fields": {f: vectors(f) for f in FIELDS}}
    path = os.path.join(ROOT, "tests", "golden", "vectors_r2.json")
    with open(path, "w") as fh:
        json.dump(data, fh, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
