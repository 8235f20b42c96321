"""The oracle has two routes through prove(): the plain per-column routines (textbook NTT, Horner, an inversion per row) and
the CPU-arm routes of oracle/fast_paths.inc (eight-column AVX2 strips interpolated once, batched inversions, tabulated powers,
block-parallel PoW search). They must give the same proof words: the fast routes exist only so that bench.py's CPU baseline
is not a strawman."""
import importlib

import numpy as np
import pytest

from common import SMALL_FRI, air_mod, field_mod, make_oracle, send_receive_system

fib = importlib.import_module("plonky3-recursion_b200.airs.fibonacci")
wl = importlib.import_module("plonky3-recursion_b200.workload")


def _both(field, fri, insts, preps, traces, pubs):
    orc = make_oracle(field, fri)
    if not orc.set_fast_paths(True):
        pytest.skip("no AVX2 on this CPU: only the plain routines exist")
    fast = orc.prove(insts, preps, traces, pubs)
    assert orc.set_fast_paths(False) is False
    plain = orc.prove(insts, preps, traces, pubs)
    orc.set_fast_paths(True)
    assert np.array_equal(fast, plain)
    cap = orc.prep_commit(insts, preps) if any(x is not None for x in preps) else None
    orc.verify(insts, cap, pubs, fast)
    return fast


@pytest.mark.parametrize("field", ["koala-bear", "baby-bear"])
def test_layer_with_lookups_prep_and_mixed_heights(field):
    F = field_mod.get_field(field)
    L = wl.synthetic_layer(F, 3, n_const=10, n_public=20, n_alu=100, n_perms=20, n_recompose=4, min_height=32)
    _both(field, SMALL_FRI, L.insts, L.preps, L.traces, L.pubs)


def test_send_receive_and_single_table_without_lookups():
    F = field_mod.get_field("koala-bear")
    insts, preps, traces, pubs = send_receive_system(F, np.random.default_rng(5))
    _both("koala-bear", SMALL_FRI, insts, preps, traces, pubs)
    t, pv = fib.trace(F.p, 7)
    inst = air_mod.build_instance("fib", fib.eval_air, F.p, 7, 2, 0, 3, air_mod.BusRegistry())
    _both("koala-bear", SMALL_FRI, [inst], [None], [t], [pv])


@pytest.mark.parametrize("fri", [dict(SMALL_FRI, log_blowup=1), dict(SMALL_FRI, log_blowup=3, max_log_arity=2),
                                 dict(SMALL_FRI, query_pow_bits=9, commit_pow_bits=3)])
def test_fri_parameter_variants(fri):
    F = field_mod.get_field("koala-bear")
    L = wl.synthetic_layer(F, 4, n_const=8, n_public=12, n_alu=40, n_perms=6, n_recompose=3, min_height=16)
    try:
        _both("koala-bear", fri, L.insts, L.preps, L.traces, L.pubs)
    except RuntimeError as e:
        if "quotient degree exceeds blowup" in str(e):
            pytest.skip("blowup 2 cannot hold the degree-3 quotient of this layer")
        raise


def test_narrow_and_odd_widths_through_the_strips():
    """Widths that are not a multiple of eight (the tail lanes of a strip), tiny heights, both row orders."""
    for field in ("koala-bear", "baby-bear"):
        F = field_mod.get_field(field)
        orc = make_oracle(field)
        if not orc.set_fast_paths(True):
            pytest.skip("no AVX2")
        rng = np.random.default_rng(11)
        for h, w, lb in ((1, 3, 2), (2, 1, 1), (8, 1, 2), (16, 7, 1), (64, 9, 2), (32, 17, 3), (128, 24, 0)):
            m = F.rand(rng, (h, w))
            plain = orc.coset_lde(m, lb)                       # rows bit-reversed
            assert np.array_equal(orc.coset_lde_strips(m, lb), plain)
            nat = orc.coset_lde_strips(m, lb, natural=True)
            bits = (h << lb).bit_length() - 1
            rev = np.array([int(format(i, f"0{bits}b")[::-1], 2) if bits else 0 for i in range(h << lb)])
            assert np.array_equal(nat, plain[rev])
