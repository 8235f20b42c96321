"""Shared helpers for the test-suite: package import, oracle construction, small synthetic table systems."""
from __future__ import annotations

import importlib

import numpy as np

pkg = importlib.import_module("plonky3-recursion_b200")
field_mod = importlib.import_module("plonky3-recursion_b200.field")
p2mod = importlib.import_module("plonky3-recursion_b200.poseidon2_params")
air_mod = importlib.import_module("plonky3-recursion_b200.air")
ws = importlib.import_module("plonky3-recursion_b200.airs.witness_send")

SMALL_FRI = dict(log_blowup=2, log_final_poly_len=2, max_log_arity=2, num_queries=6, commit_pow_bits=0, query_pow_bits=4,
                 cap_height=0)


def make_oracle(field_name="koala-bear", fri=None):
    from oracle_py import Oracle
    F = field_mod.get_field(field_name)
    return Oracle(F, p2mod.Poseidon2Params(F.field_id), dict(fri or SMALL_FRI))


def send_receive_system(F, rng, n_ops=40, log_ha=6, log_hb=5, d=4):
    """Two WitnessSend tables on one bus: table A creates n_ops witnesses (multiplicity +1), table B (2 lanes) reads
    them (multiplicity -1). Returns (instances, prep matrices, traces, public values)."""
    buses = air_mod.BusRegistry()
    vals = F.rand(rng, (n_ops, d))
    idx = np.arange(n_ops, dtype=np.uint32) * d
    A = air_mod.build_instance("const", ws.make_eval(d, 1), F.p, log_ha, d, 2, 0, buses)
    B = air_mod.build_instance("reader", ws.make_eval(d, 2), F.p, log_hb, 2 * d, 4, 0, buses)
    tA = ws.trace_to_matrix(vals, d, 1, 1 << log_ha)
    pA = ws.preprocessed_matrix(np.ones(n_ops), idx, 1, 1 << log_ha)
    tB = ws.trace_to_matrix(vals, d, 2, 1 << log_hb)
    pB = ws.preprocessed_matrix(np.full(n_ops, F.p - 1), idx, 2, 1 << log_hb)
    return [A, B], [pA, pB], [tA, tB], [None, None]
