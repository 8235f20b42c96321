#!/usr/bin/env python
"""Generate tests/golden/vectors.json: known-answer vectors produced by the CPU oracle (oracle/liboracle.so).

What these pin and what they do not: the reference (Rust, p3-* 0.6 from crates.io) cannot be built or run in this image
and ships no golden vectors for this path, so these are REGRESSION pins of the oracle's restatement, not
outputs of the reference itself ("parity unpinned" at the proof-byte level). They freeze the oracle so that
(a) an accidental change of the oracle is caught on CPU, and (b) the CUDA path is checked against committed values on the
GPU box without trusting a freshly rebuilt oracle.

Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

from common import SMALL_FRI, make_oracle, send_receive_system  # noqa: E402

wl = importlib.import_module("plonky3-recursion_b200.workload")

PERM_INPUTS = [[0] * 16, list(range(16)), [(i * 0x9E3779B1) % 0x78000001 for i in range(1, 17)]]
LAYERS = [dict(seed=7, n_const=20, n_public=30, n_alu=200, n_perms=50, n_recompose=10, min_height=32),
          dict(seed=11, n_const=5, n_public=70, n_alu=90, n_perms=17, n_recompose=3, min_height=16),
          # `recompose/coeff` table, two lanes (recompose_air.rs:175-197)
          dict(seed=41, n_const=12, n_public=50, n_alu=300, n_perms=40, n_recompose=37, min_height=32, recompose_coeff=True,
               recompose_lanes=2)]
BASE_FIB = [dict(n=1000, min_height=256), dict(n=40, min_height=16)]   # recursive_fibonacci's base circuit, extension degree 1


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint32).tobytes()).hexdigest()


def vectors(field: str) -> dict:
    orc = make_oracle(field, SMALL_FRI)
    F = orc.field
    out = {"p": F.p}
    st = np.array(PERM_INPUTS, dtype=np.uint64) % F.p
    out["poseidon2"] = {"in": st.astype(np.uint32).tolist(), "out": orc.poseidon2_permute(st.astype(np.uint32)).tolist()}
    rng = np.random.default_rng(2026)
    m = F.rand(rng, (16, 3))
    out["coset_lde"] = {"in": m.tolist(), "log_blowup": 2, "out": orc.coset_lde(m, 2).tolist()}
    mats = [F.rand(rng, (1 << lh, w)) for lh, w in ((5, 9), (4, 3), (4, 1), (2, 11))]
    out["mmcs"] = {"shapes": [[5, 9], [4, 3], [4, 1], [2, 11]], "seed": 2026, "root": orc.mmcs_commit(mats).tolist()}
    insts, preps, traces, pubs = send_receive_system(F, np.random.default_rng(1))
    proof = orc.prove(insts, preps, traces, pubs)
    out["send_receive_proof"] = {"words": int(proof.size), "sha256": sha(proof), "head": proof[:48].tolist()}
    out["layers"] = []
    for cfg in LAYERS:
        L = wl.synthetic_layer(F, cfg["seed"], **{k: v for k, v in cfg.items() if k != "seed"})
        proof = orc.prove(L.insts, L.preps, L.traces, L.pubs)
        orc.verify(L.insts, orc.prep_commit(L.insts, L.preps), L.pubs, proof)
        out["layers"].append({"cfg": cfg, "shapes": [list(s) for s in L.shapes], "words": int(proof.size), "sha256": sha(proof),
                              "prep_cap": orc.prep_commit(L.insts, L.preps).tolist(), "head": proof[:40].tolist()})
    out["base_fibonacci"] = []
    for cfg in BASE_FIB:
        L = wl.base_layer_fibonacci(F, cfg["n"], min_height=cfg["min_height"])
        proof = orc.prove(L.insts, L.preps, L.traces, L.pubs)
        orc.verify(L.insts, orc.prep_commit(L.insts, L.preps), L.pubs, proof)
        out["base_fibonacci"].append({"cfg": cfg, "shapes": [list(s) for s in L.shapes], "words": int(proof.size),
                                      "sha256": sha(proof), "prep_cap": orc.prep_commit(L.insts, L.preps).tolist(),
                                      "expected_result": int(L.traces[1][0, 0])})
    return out


if __name__ == "__main__":
    data = {"fri": SMALL_FRI, "fields": {f: vectors(f) for f in ("koala-bear", "baby-bear")}}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vectors.json")
    with open(path, "w") as f:
        json.dump(data, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")
