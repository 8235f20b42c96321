"""[P3-EXT] conventions are named switches with the same meaning in the constraint builder, the oracle and the library
(include/p3r.h p3r_conventions): every setting gives a self-consistent proof system (prove + verify), different settings give
different proofs, and the default is bit-for-bit what it was (golden vectors, generated-kernel program hashes)."""
import importlib
import itertools

import numpy as np
import pytest

from common import SMALL_FRI, field_mod, make_oracle

sym = importlib.import_module("plonky3-recursion_b200.symbolic")
wl = importlib.import_module("plonky3-recursion_b200.workload")
lib = importlib.import_module("plonky3-recursion_b200.lib")

ALL = [dict(logup_negate=n, logup_first_power=f, logup_descending=d) for n, f, d in itertools.product((0, 1), repeat=3)]


def _layer(F):
    return wl.synthetic_layer(F, 9, n_const=6, n_public=10, n_alu=60, n_perms=14, n_recompose=3, min_height=16)


@pytest.fixture
def conventions():
    saved = dict(sym.LOGUP_CONVENTIONS)
    yield sym.LOGUP_CONVENTIONS
    sym.LOGUP_CONVENTIONS.update(saved)


def test_every_convention_is_self_consistent_in_the_oracle(conventions):
    F = field_mod.get_field("koala-bear")
    orc = make_oracle("koala-bear", SMALL_FRI)
    proofs = []
    try:
        for cv in ALL:
            conventions.update(cv)
            orc.set_conventions(**cv)
            L = _layer(F)                     # instances carry the LogUp constraints of this convention
            proof = orc.prove(L.insts, L.preps, L.traces, L.pubs)
            orc.verify(L.insts, orc.prep_commit(L.insts, L.preps), L.pubs, proof)
            proofs.append(proof)
        # mismatch between builder and prover conventions is caught by the verifier
        conventions.update(ALL[0])
        L = _layer(F)
        orc.set_conventions(**ALL[1])
        bad = orc.prove(L.insts, L.preps, L.traces, L.pubs)
        with pytest.raises(RuntimeError):
            orc.verify(L.insts, orc.prep_commit(L.insts, L.preps), L.pubs, bad)
    finally:
        orc.set_conventions(**ALL[0])
    for a, b in itertools.combinations(range(len(ALL)), 2):
        assert proofs[a].size == proofs[b].size and not np.array_equal(proofs[a], proofs[b])


@pytest.mark.gpu
@pytest.mark.parametrize("cv", ALL[1:], ids=lambda c: "".join(str(v) for v in c.values()))
def test_gpu_follows_the_same_conventions(conventions, cv):
    F = field_mod.get_field("koala-bear")
    orc = make_oracle("koala-bear", SMALL_FRI)
    ctx = lib.Context("koala-bear", SMALL_FRI)
    try:
        conventions.update(cv)
        orc.set_conventions(**cv)
        ctx.set_conventions(**cv)
        L = _layer(F)
        pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
        proof = lib.BatchStarkProver(ctx).prove_all_tables(L.traces, pd, L.pubs)
        assert np.array_equal(proof, orc.prove(L.insts, L.preps, L.traces, L.pubs))
        orc.verify(L.insts, pd.preprocessed_commitment, L.pubs, proof)
        pd.close()
    finally:
        orc.set_conventions(**ALL[0])
        ctx.close()
