"""Parity against the REAL reference, when its output is available: tests/golden/ref_*.json produced by tools/ref_golden on a
machine with cargo (see tools/ref_golden/README.md). Skipped while those files do not exist (no Rust toolchain in this image)."""
import importlib
import itertools
import json
import os

import numpy as np
import pytest

from common import field_mod, make_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BASE = os.path.join(HERE, "golden", "ref_base.json")
lib = importlib.import_module("plonky3-recursion_b200.lib")
wl = importlib.import_module("plonky3-recursion_b200.workload")
sym = importlib.import_module("plonky3-recursion_b200.symbolic")

pytestmark = pytest.mark.skipif(not os.path.exists(REF_BASE), reason="no reference golden file (tools/ref_golden/README.md)")

FRI = dict(lib.DEFAULT_FRI)                       # recursive_fibonacci defaults (recursion/examples/recursive_fibonacci.rs:71-132)
CONVENTIONS = [dict(logup_negate=n, logup_first_power=f, logup_descending=d) for n, f, d in itertools.product((0, 1), repeat=3)]
SECTIONS = ["header", "degree_bits", "main_cap", "quotient/perm caps + terminals + opened values", "fri", "queries"]


def _decode(F, data):
    """The wire switches under which the bytes parse to the end of the `proof` field."""
    ok = []
    for flags in (0, lib.WIRE_CANONICAL, lib.WIRE_BARE_ROOT, lib.WIRE_CANONICAL | lib.WIRE_BARE_ROOT):
        try:
            blob, off = lib.deserialize_proof(F, FRI, data, flags)
            ok.append((flags, blob, off))
        except lib.P3RError:
            pass
    return ok


def _first_difference(a, b, n_inst):
    if a.size != b.size:
        return f"sizes differ: {a.size} vs {b.size} words"
    i = int(np.argmax(a != b))
    bounds = [5, 5 + n_inst, 5 + n_inst + 8]
    name = "later sections"
    for k, hi in enumerate(bounds):
        if i < hi:
            name = SECTIONS[k]
            break
    return f"first difference at word {i} ({name})"


def _reference_blob():
    F = field_mod.get_field("koala-bear")
    ref = json.load(open(REF_BASE))
    data = bytes.fromhex(ref["postcard_hex"])
    parses = _decode(F, data)
    assert parses, "the reference bytes parse under none of the wire switches: the field mapping in csrc/wire.cpp is off"
    # Montgomery-vs-canonical from the probe: F::TWO serialises as 2 (canonical) or as 2 * 2^32 mod p (Montgomery)
    two = int(ref["probe_koala_bear"]["two_serde"])
    canonical = two == 2
    parses = [p for p in parses if bool(p[0] & lib.WIRE_CANONICAL) == canonical]
    assert len(parses) == 1, f"ambiguous wire switches: {[p[0] for p in parses]}"
    print(f"reference wire format: flags = {parses[0][0]} (canonical = {canonical})")
    return F, parses[0][1]


def test_oracle_reproduces_the_reference_base_proof():
    F, want = _reference_blob()
    saved = dict(sym.LOGUP_CONVENTIONS)
    orc = make_oracle("koala-bear", FRI)
    report = []
    try:
        for cv in CONVENTIONS:
            sym.LOGUP_CONVENTIONS.update(cv)
            orc.set_conventions(**cv)
            L = wl.base_layer_fibonacci(F, 1000, min_height=256)
            got = orc.prove(L.insts, L.preps, L.traces, L.pubs)
            if got.size == want.size and np.array_equal(got, want):
                print(f"reference conventions: {cv}")
                return
            report.append(f"{cv}: {_first_difference(got, want, len(L.insts))}")
    finally:
        sym.LOGUP_CONVENTIONS.update(saved)
        orc.set_conventions(**CONVENTIONS[0])
    pytest.fail("no convention setting reproduces the reference proof:\n" + "\n".join(report))


@pytest.mark.gpu
def test_cuda_path_reproduces_the_reference_base_proof():
    F, want = _reference_blob()
    ctx = lib.Context("koala-bear", FRI)
    L = wl.base_layer_fibonacci(F, 1000, min_height=256)
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    got = lib.BatchStarkProver(ctx).prove_all_tables(L.traces, pd, L.pubs)
    assert got.size == want.size and np.array_equal(got, want), _first_difference(got, want, len(L.insts))
    pd.close()
    ctx.close()
