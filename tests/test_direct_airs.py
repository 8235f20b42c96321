"""The constraint bytecode the product runs (lowered from airs/alu.py, airs/poseidon2.py) against the oracle's hard-coded
evaluators written from the reference's Rust `eval` bodies (oracle/direct_airs.inc): same constraint values, position by position,
on RANDOM rows (every column uniform in the field — no structure a shared mistake could hide behind)."""
import importlib

import numpy as np
import pytest

from common import air_mod, field_mod, make_oracle, p2mod

alu = importlib.import_module("plonky3-recursion_b200.airs.alu")
poseidon2 = importlib.import_module("plonky3-recursion_b200.airs.poseidon2")


@pytest.mark.parametrize("field_name", ["koala-bear", "baby-bear"])
@pytest.mark.parametrize("d,lanes,k_max", [(4, 3, 4), (4, 2, 2), (4, 1, 2), (4, 4, 3), (4, 2, 5), (1, 1, 2), (1, 2, 3)])
def test_alu_bytecode_equals_direct_evaluator(field_name, d, lanes, k_max):
    F = field_mod.get_field(field_name)
    orc = make_oracle(field_name)
    mw, pw = alu.widths(d, lanes, k_max)
    inst = air_mod.build_instance("alu", alu.make_eval(d, lanes, k_max, F.w if d > 1 else None), F.p, 6, mw, pw, 0,
                                  air_mod.BusRegistry())
    rng = np.random.default_rng(1000 * d + 10 * lanes + k_max)
    for trial in range(6):
        local, nxt = F.rand(rng, (mw,)), F.rand(rng, (mw,))
        pl, pn = F.rand(rng, (pw,)), F.rand(rng, (pw,))
        if trial == 0:      # selectors as a real table has them: 0/1 flags instead of random values
            pl[:], pn[:] = 0, 0
            pl[alu.MULT_A] = pn[alu.MULT_A] = F.p - 1
            pl[alu.SEL_HORNER] = pn[alu.SEL_HORNER] = 1
        sel = F.rand(rng, (3,))
        want = orc.alu_eval_direct(d, lanes, k_max, local, nxt, pl, pn)
        got = orc.eval_air_rows(inst, local, nxt, pl, pn, sel)
        assert got.shape == want.shape and got.size > 0
        assert np.array_equal(got, want), f"first difference at constraint {int(np.argmax(got != want))}"


@pytest.mark.parametrize("field_name", ["koala-bear", "baby-bear"])
def test_poseidon2_bytecode_equals_direct_evaluator(field_name):
    F = field_mod.get_field(field_name)
    orc = make_oracle(field_name)
    params = p2mod.Poseidon2Params(F.field_id)
    mw, pw = poseidon2.widths(params)
    inst = air_mod.build_instance("poseidon2", poseidon2.make_eval(params), F.p, 6, mw, pw, 0, air_mod.BusRegistry())
    rng = np.random.default_rng(24)
    regs = poseidon2.sbox_registers(params)
    for trial in range(5):
        local, nxt = F.rand(rng, (mw,)), F.rand(rng, (mw,))
        pl, pn = F.rand(rng, (pw,)), F.rand(rng, (pw,))
        sel = F.rand(rng, (3,))
        if trial == 0:
            sel = np.array([0, 0, 1], dtype=np.uint32)
        want = orc.poseidon2_eval_direct(regs, local, nxt, pl, pn, int(sel[2]))
        got = orc.eval_air_rows(inst, local, nxt, pl, pn, sel)
        assert got.shape == want.shape and got.size == 1 + 16 + 16 + 1 + 8 * 16 * (1 + regs) + params.rounds_p * (1 + regs)
        assert np.array_equal(got, want), f"first difference at constraint {int(np.argmax(got != want))}"


def test_direct_evaluators_vanish_on_a_valid_table():
    """And the hard-coded evaluators accept what the table builders produce: every constraint is zero on every row pair of a
    valid synthetic layer (ALU with packed Horner chains, Poseidon2 with sponge and Merkle chains)."""
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    F = field_mod.get_field("koala-bear")
    orc = make_oracle("koala-bear")
    L = wl.synthetic_layer(F, 3, n_const=8, n_public=12, n_alu=150, n_perms=40, n_recompose=3, min_height=16)
    ta, pa = L.traces[2], L.preps[2]
    n = ta.shape[0]
    for r in range(n):
        c = orc.alu_eval_direct(4, 3, 4, ta[r], ta[(r + 1) % n], pa[r], pa[(r + 1) % n])
        assert not c.any(), f"ALU row {r}: constraint {int(np.argmax(c != 0))}"
    tp, pp = L.traces[3], L.preps[3]
    n = tp.shape[0]
    for r in range(n):
        c = orc.poseidon2_eval_direct(0, tp[r], tp[(r + 1) % n], pp[r], pp[(r + 1) % n], int(r != n - 1))
        assert not c.any(), f"Poseidon2 row {r}: constraint {int(np.argmax(c != 0))}"
