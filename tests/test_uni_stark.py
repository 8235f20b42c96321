"""Uni-STARK mode: one wide table proved with p3-uni-stark's transcript head
(recursion/src/types/challenges.rs:44-54,100-140), same kernels / same oracle phases as the batch prover."""
import importlib

import numpy as np
import pytest

from common import SMALL_FRI, air_mod, field_mod, make_oracle

lib = importlib.import_module("plonky3-recursion_b200.lib")
wide = importlib.import_module("plonky3-recursion_b200.airs.wide")


def _instance(F, width, log_n):
    inst = air_mod.build_instance("wide", wide.make_eval(width), F.p, log_n, width, 0, wide.N_PUBLIC, air_mod.BusRegistry())
    assert inst.uses_next_row and inst.log_quotient_chunks == 1        # degree 3 -> two quotient chunks
    t, pubs = wide.trace(F.p, width, log_n)
    return inst, t, pubs


@pytest.mark.parametrize("field_name", ["koala-bear", "baby-bear"])
def test_oracle_uni_stark_round_trip_and_differs_from_batch(field_name):
    F = field_mod.get_field(field_name)
    orc = make_oracle(field_name, SMALL_FRI)
    inst, t, pubs = _instance(F, 41, 5)
    assert orc.check_constraints(inst, None, t, pubs) is None
    batch = orc.prove([inst], [None], [t], [pubs])
    try:
        orc.set_uni_stark(True)
        uni = orc.prove([inst], [None], [t], [pubs])
        orc.verify([inst], None, [pubs], uni)
        with pytest.raises(RuntimeError):
            orc.verify([inst], None, [pubs], batch)               # a batch-transcript proof is not a uni-stark proof
        bad = pubs.copy()
        bad[2] = (int(bad[2]) + 1) % F.p
        with pytest.raises(RuntimeError):
            orc.verify([inst], None, [bad], uni)
    finally:
        orc.set_uni_stark(False)
    assert uni.size == batch.size and not np.array_equal(uni, batch)
    n = 1                                                          # identical up to the main commitment (same trace, same LDE)
    assert np.array_equal(uni[: 5 + n + 8], batch[: 5 + n + 8]) and not np.array_equal(uni[5 + n + 8: 5 + n + 16], batch[5 + n + 8: 5 + n + 16])


@pytest.mark.gpu
@pytest.mark.parametrize("field_name,width,log_n", [("koala-bear", 41, 5), ("baby-bear", 100, 7), ("koala-bear", 650, 9)])
def test_gpu_uni_stark_bit_identical(field_name, width, log_n):
    F = field_mod.get_field(field_name)
    orc = make_oracle(field_name, SMALL_FRI)
    ctx = lib.Context(field_name, SMALL_FRI)
    inst, t, pubs = _instance(F, width, log_n)
    try:
        orc.set_uni_stark(True)
        ctx.set_uni_stark(True)
        pd = lib.ProverData.from_airs_and_degrees(ctx, [inst], [None])
        proof = lib.BatchStarkProver(ctx).prove_all_tables([t], pd, [pubs])
        assert np.array_equal(proof, orc.prove([inst], [None], [t], [pubs]))
        orc.verify([inst], None, [pubs], proof)
        pd.close()
    finally:
        orc.set_uni_stark(False)
        ctx.close()


@pytest.mark.gpu
def test_gpu_keccak_shaped_base_layer_is_accepted_by_the_oracle_verifier():
    """The recursive_keccak base-layer shape: 2 600 columns x 4 096 rows (recursive_keccak.rs:22-24,513-517), example FRI
    parameters: the oracle's verifier accepts the GPU proof and the oracle's prover produces the same words."""
    F = field_mod.get_field("koala-bear")
    orc = make_oracle("koala-bear", lib.DEFAULT_FRI)
    ctx = lib.Context("koala-bear", lib.DEFAULT_FRI)
    inst, t, pubs = _instance(F, 2600, 12)
    try:
        orc.set_uni_stark(True)
        ctx.set_uni_stark(True)
        pd = lib.ProverData.from_airs_and_degrees(ctx, [inst], [None])
        prover = lib.BatchStarkProver(ctx)
        proof = prover.prove_all_tables([t], pd, [pubs])
        orc.verify([inst], None, [pubs], proof)
        assert np.array_equal(proof, orc.prove([inst], [None], [t], [pubs]))
        bad = proof.copy()
        bad[proof.size // 2] = (int(bad[proof.size // 2]) + 1) % F.p
        with pytest.raises(RuntimeError):
            orc.verify([inst], None, [pubs], bad)
        pd.close()
    finally:
        orc.set_uni_stark(False)
        ctx.close()
