"""GPU parity tests: every result of the CUDA path must be bit-identical to the CPU oracle on the same seeded inputs.
All calls go through the C ABI (lib.py -> libp3r_b200.so)."""
import importlib

import numpy as np
import pytest

from common import SMALL_FRI, field_mod, make_oracle, send_receive_system

lib = importlib.import_module("plonky3-recursion_b200.lib")

pytestmark = pytest.mark.gpu

FIELDS = ["koala-bear", "baby-bear"]


@pytest.fixture(scope="module", params=FIELDS)
def pair(request):
    ctx = lib.Context(request.param, SMALL_FRI)
    orc = make_oracle(request.param, SMALL_FRI)
    yield ctx, orc
    ctx.close()


def test_poseidon2_permutation(pair):
    ctx, orc = pair
    rng = np.random.default_rng(7)
    st = ctx.field.rand(rng, (257, 16))
    st[0] = 0
    st[1] = ctx.field.p - 1
    assert np.array_equal(ctx.poseidon2_permute(st), orc.poseidon2_permute(st))


@pytest.mark.parametrize("log_n,width,log_blowup", [(3, 1, 1), (5, 3, 2), (8, 17, 2), (10, 5, 3), (13, 4, 1), (14, 3, 2), (16, 2, 1)])
def test_coset_lde(pair, log_n, width, log_blowup):
    ctx, orc = pair
    rng = np.random.default_rng(100 + log_n)
    m = ctx.field.rand(rng, (1 << log_n, width))
    got = ctx.coset_lde(m, log_blowup)
    want = orc.coset_lde(m, log_blowup)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("shapes", [[(6, 5)], [(8, 9), (8, 16)], [(9, 3), (7, 20), (7, 1), (4, 11)], [(10, 8), (9, 8), (3, 8)]])
def test_mmcs_commit_mixed_heights(pair, shapes):
    ctx, orc = pair
    rng = np.random.default_rng(5)
    mats = [ctx.field.rand(rng, (1 << lh, w)) for lh, w in shapes]
    assert np.array_equal(ctx.mmcs_commit(mats), orc.mmcs_commit(mats))


def test_grind_smallest_witness(pair):
    ctx, orc = pair
    rng = np.random.default_rng(11)
    F = ctx.field
    state = F.to_monty(F.rand(rng, 16))
    for n_pending in (0, 3, 7):
        pending = F.to_monty(F.rand(rng, n_pending))
        assert ctx.grind(state, pending, 9) == orc.grind(state, pending, 9)


def test_full_proof_bit_identical_and_verifies(pair):
    ctx, orc = pair
    rng = np.random.default_rng(3)
    insts, preps, traces, pubs = send_receive_system(ctx.field, rng)
    pd = lib.ProverData.from_airs_and_degrees(ctx, insts, preps)
    proof = lib.BatchStarkProver(ctx).prove_all_tables(traces, pd, pubs)
    want = orc.prove(insts, preps, traces, pubs)
    assert np.array_equal(pd.preprocessed_commitment, orc.prep_commit(insts, preps))
    assert proof.size == want.size
    diff = np.nonzero(proof != want)[0]
    assert diff.size == 0, f"first differing word {diff[:5]} of {proof.size}"
    orc.verify(insts, pd.preprocessed_commitment, pubs, proof)
    pd.close()


@pytest.mark.parametrize("seed", [1, 2])
def test_recursion_layer_tables_bit_identical(pair, seed):
    """Const/Public/ALU(3 lanes, k=4)/Poseidon2/Recompose with the WitnessChecks LogUp bus, mixed heights."""
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    ctx, orc = pair
    L = wl.synthetic_layer(ctx.field, seed, n_const=20, n_public=70, n_alu=400, n_perms=90, n_recompose=10, min_height=32)
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    proof = lib.BatchStarkProver(ctx).prove_all_tables(L.traces, pd, L.pubs)
    want = orc.prove(L.insts, L.preps, L.traces, L.pubs)
    assert proof.size == want.size
    diff = np.nonzero(proof != want)[0]
    assert diff.size == 0, f"first differing word {diff[:5]} of {proof.size}"
    orc.verify(L.insts, pd.preprocessed_commitment, L.pubs, proof)
    # tampering with any opened value or commitment must be rejected
    bad = proof.copy()
    bad[30] ^= 1
    with pytest.raises(RuntimeError):
        orc.verify(L.insts, pd.preprocessed_commitment, L.pubs, bad)
    pd.close()
