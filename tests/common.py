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


class PyChallenger:
    """DuplexChallenger<F, Perm, 16, 8> in plain Python/numpy, the transcript a Rust host would own when it
    drives the phase-stepped C ABI. Values are canonical."""

    def __init__(self, F):
        self.F = F
        self.prm = p2mod.Poseidon2Params(F.field_id)
        self.state = [0] * 16
        self.inp, self.out = [], []

    def _duplex(self):
        n = len(self.inp)
        for i, v in enumerate(self.inp):
            self.state[i] = v
        if n:
            for i in range(n, 8):
                self.state[i] = 0
            self.state[8] = (self.state[8] + n) % self.F.p
        self.inp = []
        self.state = [int(x) for x in self.prm.permute(np.array(self.state, dtype=np.uint64)[None])[0]]
        self.out = list(self.state[:8])

    def observe(self, v):
        self.out = []
        self.inp.append(int(v) % self.F.p)
        if len(self.inp) == 8:
            self._duplex()

    def observe_many(self, vs):
        for v in vs:
            self.observe(v)

    def observe_lifted(self, v):
        self.observe_many([v, 0, 0, 0])

    def sample(self):
        if self.inp or not self.out:
            self._duplex()
        return self.out.pop()

    def sample_ext(self):
        return [self.sample() for _ in range(4)]

    def sample_bits(self, bits):
        return self.sample() & ((1 << bits) - 1)
