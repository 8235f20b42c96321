"""CPU oracle prover <-> verifier: round trips over the proof shapes and FRI parameter variants the reference exercises
(circuit-prover/src/batch_stark_prover/tests.rs, recursion/tests/fri.rs), plus rejection of tampered proofs and of
invalid witnesses. These run without a GPU; the GPU suite repeats the same systems through the C ABI bit-for-bit."""
import importlib

import numpy as np
import pytest

from common import SMALL_FRI, air_mod, field_mod, make_oracle, send_receive_system, ws

fib = importlib.import_module("plonky3-recursion_b200.airs.fibonacci")
wl = importlib.import_module("plonky3-recursion_b200.workload")


def fib_system(F, log_n=6):
    t, pubs = fib.trace(F.p, log_n)
    inst = air_mod.build_instance("fib", fib.eval_air, F.p, log_n, 2, 0, 3, air_mod.BusRegistry())
    return [inst], [None], [t], [pubs]


@pytest.mark.parametrize("field", ["koala-bear", "baby-bear"])
def test_send_receive_roundtrip_and_tamper(field):
    orc = make_oracle(field)
    rng = np.random.default_rng(1)
    insts, preps, traces, pubs = send_receive_system(orc.field, rng)
    proof = orc.prove(insts, preps, traces, pubs)
    cap = orc.prep_commit(insts, preps)
    orc.verify(insts, cap, pubs, proof)
    assert np.array_equal(proof, orc.prove(insts, preps, traces, pubs))  # deterministic (smallest PoW witness)
    # header, commitments, terminals, opened values, FRI caps, final polynomial, PoW witness, query data
    for pos in (1, 9, 17, 30, 33, 60, 200, proof.size // 2, proof.size - 5):
        bad = proof.copy()
        bad[pos] = (int(bad[pos]) + 1) % orc.field.p
        with pytest.raises(RuntimeError):
            orc.verify(insts, cap, pubs, bad)
    with pytest.raises(RuntimeError):
        orc.verify(insts, cap, pubs, proof[:-1])
    wrong_cap = cap.copy()
    wrong_cap[0] ^= 1
    with pytest.raises(RuntimeError):
        orc.verify(insts, wrong_cap, pubs, proof)


def test_unbalanced_bus_is_rejected():
    orc = make_oracle("koala-bear")
    rng = np.random.default_rng(2)
    insts, preps, traces, pubs = send_receive_system(orc.field, rng)
    preps[0] = preps[0].copy()
    preps[0][3, 0] = 2  # creator claims two reads, only one reader exists
    proof = orc.prove(insts, preps, traces, pubs)
    with pytest.raises(RuntimeError, match="terminals"):
        orc.verify(insts, orc.prep_commit(insts, preps), pubs, proof)
    # a reader looking up a value nobody created is also caught
    insts, preps, traces, pubs = send_receive_system(orc.field, rng)
    traces[1] = traces[1].copy()
    traces[1][0, 0] = (int(traces[1][0, 0]) + 1) % orc.field.p
    proof = orc.prove(insts, preps, traces, pubs)
    with pytest.raises(RuntimeError, match="terminals"):
        orc.verify(insts, orc.prep_commit(insts, preps), pubs, proof)


@pytest.mark.parametrize("field", ["koala-bear", "baby-bear"])
def test_fibonacci_public_values_no_lookups_no_prep(field):
    orc = make_oracle(field)
    insts, preps, traces, pubs = fib_system(orc.field)
    assert insts[0].log_quotient_chunks == 0 and not insts[0].lookups and insts[0].uses_next_row
    proof = orc.prove(insts, preps, traces, pubs)
    assert proof[2] == 0 and proof[3] == 0  # no permutation / preprocessed commitments in the proof
    orc.verify(insts, None, pubs, proof)
    wrong = [pubs[0].copy()]
    wrong[0][2] = (int(wrong[0][2]) + 1) % orc.field.p
    with pytest.raises(RuntimeError):
        orc.verify(insts, None, wrong, proof)
    # an invalid witness yields a proof the verifier rejects: constraints(zeta)/Z_H(zeta) != Q(zeta)
    bad = traces[0].copy()
    bad[10, 1] = (int(bad[10, 1]) + 1) % orc.field.p
    bad_proof = orc.prove(insts, preps, [bad], pubs)
    with pytest.raises(RuntimeError, match="constraint/quotient"):
        orc.verify(insts, None, pubs, bad_proof)


@pytest.mark.parametrize("fri", [
    dict(SMALL_FRI, cap_height=2),
    dict(SMALL_FRI, max_log_arity=1),
    dict(SMALL_FRI, max_log_arity=3, log_final_poly_len=0),
    dict(SMALL_FRI, commit_pow_bits=3, query_pow_bits=0),
    dict(SMALL_FRI, log_blowup=1, num_queries=10),
    dict(SMALL_FRI, log_blowup=3, log_final_poly_len=1),
])
def test_fri_parameter_variants(fri):
    orc = make_oracle("koala-bear", fri)
    L = wl.synthetic_layer(orc.field, 4, n_const=10, n_public=20, n_alu=120, n_perms=30, n_recompose=5, min_height=64)
    proof = orc.prove(L.insts, L.preps, L.traces, L.pubs)
    orc.verify(L.insts, orc.prep_commit(L.insts, L.preps), L.pubs, proof)
    n_inst = len(L.insts)
    cap_words = 8 << fri["cap_height"]
    assert proof[1] == n_inst and proof[4] == cap_words


@pytest.mark.parametrize("field,lanes", [("koala-bear", 1), ("baby-bear", 2)])
def test_recompose_with_coefficient_lookups(field, lanes):
    """Layer whose Recompose table is the `recompose/coeff` variant (recompose_air.rs:175-197): the bus balances only if
    every coefficient receive (idx, v_i, 0, 0, 0) is matched, so a wrong coefficient multiplicity must be rejected."""
    orc = make_oracle(field)
    L = wl.synthetic_layer(orc.field, 8, n_const=10, n_public=20, n_alu=150, n_perms=20, n_recompose=7, min_height=32,
                           recompose_coeff=True, recompose_lanes=lanes)
    proof = orc.prove(L.insts, L.preps, L.traces, L.pubs)
    orc.verify(L.insts, orc.prep_commit(L.insts, L.preps), L.pubs, proof)
    preps = [None if m is None else m.copy() for m in L.preps]
    row = np.nonzero(preps[-1][:, 3])[0][0]
    preps[-1][row, 3] += 1          # coeff_0_mult of a hint-output operation
    bad = orc.prove(L.insts, preps, L.traces, L.pubs)
    with pytest.raises(RuntimeError):
        orc.verify(L.insts, orc.prep_commit(L.insts, preps), L.pubs, bad)


@pytest.mark.parametrize("field", ["koala-bear", "baby-bear"])
def test_base_layer_fibonacci_circuit(field):
    """BASELINE configs[0]'s base circuit (recursive_fibonacci --n 1000, extension degree 1, TablePacking::new(1, 1)):
    shapes of air/shape_golden.rs (ALU 1024 x 7 / 20, Const and Public 256 x 1), F(1000) as the public witness that the last
    ADD reads back (`connect`), proof accepted; a wrong expected_result unbalances the bus and is rejected."""
    orc = make_oracle(field)
    F = orc.field
    L = wl.base_layer_fibonacci(F, 1000)
    assert L.shapes == [("const", 256, 1, 2), ("public", 256, 1, 2), ("alu", 1024, 7, 20)]
    a, b = 0, 1
    for _ in range(2, 1001):
        a, b = b, (a + b) % F.p
    assert int(L.traces[1][0, 0]) == b and int(L.traces[2][998, 3]) == b and not L.traces[2][999:].any()
    cap = orc.prep_commit(L.insts, L.preps)
    proof = orc.prove(L.insts, L.preps, L.traces, L.pubs)
    orc.verify(L.insts, cap, L.pubs, proof)
    traces = [t.copy() for t in L.traces]
    traces[1][0, 0] = (b + 1) % F.p
    with pytest.raises(RuntimeError):
        orc.verify(L.insts, cap, L.pubs, orc.prove(L.insts, L.preps, traces, L.pubs))


def test_layer_with_public_values_and_mixed_heights():
    """Fibonacci (public values, no lookups) next to the 5 recursion tables: instances with and without lookups /
    preprocessed columns / next-row openings in one batch."""
    orc = make_oracle("baby-bear")
    F = orc.field
    L = wl.synthetic_layer(F, 6, n_const=10, n_public=20, n_alu=100, n_perms=30, n_recompose=5, min_height=32)
    fi, fp, ft, fpub = fib_system(F, 7)
    insts, preps, traces, pubs = L.insts + fi, L.preps + fp, L.traces + ft, L.pubs + fpub
    proof = orc.prove(insts, preps, traces, pubs)
    orc.verify(insts, orc.prep_commit(insts, preps), pubs, proof)


@pytest.mark.parametrize("public_lanes,alu_lanes,horner_k", [(1, 1, 2), (2, 2, 4), (2, 4, 3), (4, 3, 2)])
def test_table_packing_variants(public_lanes, alu_lanes, horner_k):
    """`TablePacking::new(public_lanes, alu_lanes)` with different packed-Horner depths (alu_air.rs:22-58): widths follow
    `widths()`, every AIR holds on its trace, the bus balances and the proof verifies."""
    orc = make_oracle("koala-bear")
    L = wl.synthetic_layer(orc.field, 13, n_const=10, n_public=33, n_alu=260, n_perms=20, n_recompose=5, min_height=16,
                           public_lanes=public_lanes, alu_lanes=alu_lanes, horner_k=horner_k)
    alu_mod = importlib.import_module("plonky3-recursion_b200.airs.alu")
    assert (L.insts[2].main_width, L.insts[2].prep_width) == alu_mod.widths(4, alu_lanes, horner_k)
    assert (L.insts[1].main_width, L.insts[1].prep_width) == (4 * public_lanes, 2 * public_lanes)
    assert len(L.insts[2].interactions) == 4 * alu_lanes + 2 * (horner_k - 1)
    for s, pm, tr in zip(L.insts, L.preps, L.traces):
        assert orc.check_constraints(s, pm, tr, None) is None, s.name
    proof = orc.prove(L.insts, L.preps, L.traces, L.pubs)
    orc.verify(L.insts, orc.prep_commit(L.insts, L.preps), L.pubs, proof)
