"""Poseidon2 width 24 (BASELINE north_star (3); circuit/src/ops/poseidon2_perm/config.rs:77-86,124-133): permutation, the
PaddingFreeSponge<Perm24, 24, 16, 8> leaf hasher, and whole proofs in the "width-24 hash / width-16 compress" configuration."""
import importlib

import numpy as np
import pytest

from common import SMALL_FRI, field_mod, make_oracle, p2mod

lib = importlib.import_module("plonky3-recursion_b200.lib")
wl = importlib.import_module("plonky3-recursion_b200.workload")
FIELDS = ["koala-bear", "baby-bear"]


def test_round_constants_match_the_p3_tables_as_recalled():
    """The Grain LFSR with t = 24 reproduces the first constants of p3-baby-bear's BABYBEAR_RC24_EXTERNAL_INITIAL as recalled
    (0x0fa20c37, 0x0795bb97, 0x12c60b9c): the same generator that reproduces the width-16 tables (tests/test_oracle_kat.py)."""
    bb = p2mod.Poseidon2Params(field_mod.BABYBEAR, 24)
    assert [int(x) for x in bb.external_rc[:3]] == [0x0FA20C37, 0x0795BB97, 0x12C60B9C]
    assert bb.rounds_p == 21 and bb.external_rc.size == 8 * 24 and bb.internal_diag.size == 24
    kb = p2mod.Poseidon2Params(field_mod.KOALABEAR, 24)
    assert kb.rounds_p == 23 and kb.sbox_degree == 3


@pytest.mark.parametrize("field_name", FIELDS)
@pytest.mark.parametrize("width", [16, 24])
def test_oracle_generic_permutation_matches_numpy(field_name, width):
    F = field_mod.get_field(field_name)
    prm = p2mod.Poseidon2Params(F.field_id, width)
    orc = make_oracle(field_name)
    st = F.rand(np.random.default_rng(width), (19, width))
    st[0] = 0
    got = orc.poseidon2_permute_w(prm, st)
    assert np.array_equal(got, prm.permute(st))
    if width == 16:
        assert np.array_equal(got, orc.poseidon2_permute(st))      # the specialised width-16 code path agrees


@pytest.mark.parametrize("field_name", FIELDS)
def test_oracle_width24_leaf_hashing_changes_the_commitment_and_stays_consistent(field_name):
    F = field_mod.get_field(field_name)
    orc = make_oracle(field_name, SMALL_FRI)
    prm = p2mod.Poseidon2Params(F.field_id, 24)
    rng = np.random.default_rng(3)
    mats = [F.rand(rng, (64, 37)), F.rand(rng, (16, 5))]
    c16 = orc.mmcs_commit(mats).copy()
    try:
        orc.set_leaf_hasher(prm)
        c24 = orc.mmcs_commit(mats).copy()
        orc.mmcs_open_verify(mats, 11)
        # single matrix of one row group: the cap of a height-1 "tree" is the sponge digest itself
        one = F.rand(rng, (1, 37))
        assert np.array_equal(F.from_monty(orc.mmcs_commit([one])), prm.sponge(one, 16)[0])
        L = wl.synthetic_layer(F, 4, n_const=6, n_public=9, n_alu=40, n_perms=10, n_recompose=3, min_height=16)
        proof = orc.prove(L.insts, L.preps, L.traces, L.pubs)
        orc.verify(L.insts, orc.prep_commit(L.insts, L.preps), L.pubs, proof)
    finally:
        orc.set_leaf_hasher(None)
    assert not np.array_equal(c16, c24)
    assert np.array_equal(orc.mmcs_commit(mats), c16)


@pytest.mark.gpu
@pytest.mark.parametrize("field_name", FIELDS)
def test_gpu_width24_permutation_and_proofs(field_name):
    F = field_mod.get_field(field_name)
    orc = make_oracle(field_name, SMALL_FRI)
    ctx = lib.Context(field_name, SMALL_FRI)
    prm = p2mod.Poseidon2Params(F.field_id, 24)
    st = F.rand(np.random.default_rng(5), (300, 24))
    assert np.array_equal(ctx.poseidon2_permute_w(prm, st), prm.permute(st))
    st16 = F.rand(np.random.default_rng(6), (70, 16))
    assert np.array_equal(ctx.poseidon2_permute_w(p2mod.Poseidon2Params(F.field_id, 16), st16), orc.poseidon2_permute(st16))
    rng = np.random.default_rng(8)
    mats = [F.rand(rng, (1 << lh, w)) for lh, w in [(10, 3), (9, 70), (8, 170), (10, 33), (4, 17)]]
    try:
        orc.set_leaf_hasher(prm)
        ctx.set_leaf_hasher(prm)
        assert np.array_equal(ctx.mmcs_commit(mats), orc.mmcs_commit(mats))
        for seed, big in ((4, False), (5, True)):
            L = wl.synthetic_layer(F, seed, n_const=6, n_public=20 if big else 9, n_alu=600 if big else 40,
                                   n_perms=100 if big else 10, n_recompose=3, min_height=16)
            pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
            proof = lib.BatchStarkProver(ctx).prove_all_tables(L.traces, pd, L.pubs)
            assert np.array_equal(proof, orc.prove(L.insts, L.preps, L.traces, L.pubs))
            orc.verify(L.insts, pd.preprocessed_commitment, L.pubs, proof)
            pd.close()
        ctx.set_leaf_hasher(None)
        orc.set_leaf_hasher(None)
        assert np.array_equal(ctx.mmcs_commit(mats), orc.mmcs_commit(mats))
    finally:
        orc.set_leaf_hasher(None)
        ctx.close()
