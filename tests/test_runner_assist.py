"""Runner assist: p3r_poseidon2_run_chains executes the Poseidon2 rows of the circuit runner as chains
on the device (circuit/src/ops/poseidon_perm/executor.rs:924-975). Checked against the host path the synthetic workload uses
(one dependent permutation per row) and against the table builder."""
import importlib

import numpy as np
import pytest

from common import field_mod, make_oracle

lib = importlib.import_module("plonky3-recursion_b200.lib")
wl = importlib.import_module("plonky3-recursion_b200.workload")

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("field_name", ["koala-bear", "baby-bear"])
def test_chains_match_the_host_runner(field_name):
    F = field_mod.get_field(field_name)
    ctx = lib.Context(field_name)
    orc = make_oracle(field_name)
    rng = np.random.default_rng(12)
    W = wl.Witnesses()
    src = [W.new(wl._rand_ext(F, rng)) for _ in range(50)]
    pubs = []
    npo = wl.NpoTables(F, rng, W, src, lambda v: pubs.append(W.new(v)) or pubs[-1], n_perms=700, n_recompose=0)
    rows = npo.rows
    n = len(rows)
    new_start = np.array([r["new_start"] for r in rows], dtype=np.uint8)
    merkle = np.array([r["merkle_path"] for r in rows], dtype=np.uint8)
    bit = np.array([r["mmcs_bit"] for r in rows], dtype=np.uint8)
    want_in = np.array([r["inputs"] for r in rows], dtype=np.uint32)
    mask = np.zeros(n, dtype=np.uint8)
    values = np.zeros((n, 16), dtype=np.uint32)
    for k, r in enumerate(rows):
        if r["merkle_path"]:
            # the sibling is private data: it sits in the half the running digest does NOT occupy after the swap
            values[k, 8:16] = want_in[k, 0:8] if r["mmcs_bit"] else want_in[k, 8:16]
        else:
            for l in range(4):
                if r["in_ctl"][l]:
                    mask[k] |= 1 << l
                    values[k, 4 * l:4 * l + 4] = want_in[k, 4 * l:4 * l + 4]   # chained limbs stay zero: the kernel must supply them
    got_in, got_out = ctx.poseidon2_run_chains(new_start, merkle, bit, mask, values)
    assert np.array_equal(got_in, want_in)
    assert np.array_equal(got_out, orc.poseidon2_permute(want_in))
    # chains really are chained: the continuation rows' inputs contain the previous row's output
    k = next(i for i in range(1, n) if not new_start[i] and not merkle[i])
    assert np.array_equal(got_in[k, 8:16], got_out[k - 1, 8:16])
    ctx.close()
