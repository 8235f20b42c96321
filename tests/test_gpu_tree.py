"""The aggregation tree on one GPU through the same code bench.py times (bench.Lane + aggregation.TreeExecutor): real proofs,
child checksums folded into the parents' Public tables, device-resident and host-buffer modes."""
import importlib

import numpy as np
import pytest

import bench
from common import SMALL_FRI, field_mod, make_oracle

lib = importlib.import_module("plonky3-recursion_b200.lib")
agg = importlib.import_module("plonky3-recursion_b200.aggregation")
wl = importlib.import_module("plonky3-recursion_b200.workload")

pytestmark = pytest.mark.gpu


def _shapes(F):
    return {"leaf": wl.base_layer_fibonacci(F, 50, min_height=32),
            "l1": wl.synthetic_layer(F, 2, n_const=10, n_public=40, n_alu=150, n_perms=20, n_recompose=5, min_height=32,
                                     public_lanes=2, alu_lanes=2, horner_k=2),
            "node": wl.synthetic_layer(F, 1, n_const=12, n_public=40, n_alu=200, n_perms=40, n_recompose=6, min_height=32)}


def test_tree_roots_depend_on_every_leaf_and_match_a_serial_run():
    F = field_mod.get_field("koala-bear")
    shapes = _shapes(F)
    lanes = [bench.Lane(lib, "koala-bear", SMALL_FRI, 0, shapes) for _ in range(3)]
    kind = lambda lvl: "leaf" if lvl == 0 else ("l1" if lvl == 1 else "node")
    host = {"on": False}
    swap = {"on": False}

    def leaf(k, t, i):
        ident = np.zeros(bench.PATCH_WORDS, dtype=np.uint32)
        ident[0], ident[1] = t, i
        return lanes[k].prove("leaf", ident, host["on"])

    def node(k, t, nd, left, right):
        if swap["on"] and nd.level == 2 and nd.index == 1:
            left, right = right, left                        # a mis-routed pair of child proofs
        patch = np.concatenate([bench.proof_checksum(left, F.p), bench.proof_checksum(right, F.p)])
        return lanes[k].prove(kind(nd.level), patch, host["on"])

    ex = agg.TreeExecutor(0, 1, 8, 3, leaf, node)
    out = ex.run(3)
    assert out["proved"] == {0: 24, 1: 12, 2: 6, 3: 3}
    # serial recomputation on one lane, in the reference's loop order (recursive_aggregation.rs:676-704)
    for t in range(3):
        cur = [leaf(0, t, i) for i in range(8)]
        lvl = 0
        while len(cur) > 1:
            lvl += 1
            cur = [node(0, t, agg.Node(lvl, i), cur[2 * i], cur[2 * i + 1]) for i in range(len(cur) // 2)]
        assert np.array_equal(cur[0], out["roots"][t])
    assert not np.array_equal(out["roots"][0], out["roots"][1])          # the leaf identity reaches the root
    # the root is a valid proof of the node shape for the oracle's verifier
    orc = make_oracle("koala-bear", SMALL_FRI)
    pd = lanes[0].kind["node"][0]
    orc.verify(shapes["node"].insts, pd.preprocessed_commitment, shapes["node"].pubs, out["roots"][2])
    # host-buffer mode (operation lists / matrices uploaded per proof) gives the same bytes
    host["on"] = True
    out_h = agg.TreeExecutor(0, 1, 8, 3, leaf, node).run(2)
    assert np.array_equal(out_h["roots"][0], out["roots"][0]) and np.array_equal(out_h["roots"][1], out["roots"][1])
    # swapping two child proofs below the root changes the root
    swap["on"] = True
    out_s = agg.TreeExecutor(0, 1, 8, 3, leaf, node).run(1)
    assert not np.array_equal(out_s["roots"][0], out["roots"][0])
    for ln in lanes:
        ln.close()


def test_write_rows_matches_a_fresh_upload():
    """p3r_traces_write_rows on resident traces == uploading the patched matrix; the proof equals the oracle's on that matrix."""
    F = field_mod.get_field("koala-bear")
    L = wl.synthetic_layer(F, 5, n_const=10, n_public=20, n_alu=100, n_perms=20, n_recompose=4, min_height=32)
    ctx = lib.Context("koala-bear", SMALL_FRI)
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    tb = lib.TraceBatch(ctx, L.traces, L.pubs).upload(pd)
    rows = F.rand(np.random.default_rng(3), (4, 4))
    H = L.traces[1].shape[0]
    tb.write_rows(pd, 1, H - 4, rows)
    got = tb.download(pd, 1)
    want = L.traces[1].copy()
    want[H - 4:] = rows
    assert np.array_equal(got, want)
    proof = lib.BatchStarkProver(ctx).prove_resident(tb, pd)
    traces = list(L.traces)
    traces[1] = want
    orc = make_oracle("koala-bear", SMALL_FRI)
    assert np.array_equal(proof, orc.prove(L.insts, L.preps, traces, L.pubs))
    with pytest.raises(lib.P3RError):
        tb.write_rows(pd, 1, H - 2, rows)                     # past the end of the table
    tb.close()
    pd.close()
    ctx.close()


def test_block_order_and_priority_streams_give_the_same_roots():
    """The task order and the CUDA stream priority are scheduling only: trees proved in block order, with the leaf and level-1
    shapes on high-priority contexts (p3r_ctx_set_stream_priority), have the same root proofs as the wave-order run."""
    F = field_mod.get_field("koala-bear")
    shapes = _shapes(F)
    kind = lambda lvl: "leaf" if lvl == 0 else ("l1" if lvl == 1 else "node")
    roots = []
    for order, prio in (("wave", ()), ("block", ("leaf", "l1"))):
        lanes = [bench.Lane(lib, "koala-bear", SMALL_FRI, 0, shapes, prio) for _ in range(3)]
        assert all(len(ln.ctxs) == (2 if prio else 1) for ln in lanes)

        def leaf(k, t, i):
            ident = np.zeros(bench.PATCH_WORDS, dtype=np.uint32)
            ident[0], ident[1] = t, i
            return lanes[k].prove("leaf", ident, False)

        def node(k, t, nd, left, right):
            patch = np.concatenate([bench.proof_checksum(left, F.p), bench.proof_checksum(right, F.p)])
            return lanes[k].prove(kind(nd.level), patch, False)

        out = agg.TreeExecutor(0, 1, 8, 3, leaf, node, order=order, skew=2).run(5)
        roots.append(out["roots"])
        for ln in lanes:
            ln.close()
    for t in range(5):
        assert np.array_equal(roots[0][t], roots[1][t])


def test_stream_priority_can_be_switched_between_proofs():
    F = field_mod.get_field("koala-bear")
    L = wl.synthetic_layer(F, 7, n_const=10, n_public=20, n_alu=100, n_perms=20, n_recompose=4, min_height=32)
    ctx = lib.Context("koala-bear", SMALL_FRI)
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    tb = lib.TraceBatch(ctx, L.traces, L.pubs).upload(pd)
    prover = lib.BatchStarkProver(ctx)
    a = prover.prove_resident(tb, pd)
    ctx.set_stream_priority(True)
    b = prover.prove_resident(tb, pd)
    ctx.set_stream_priority(False)
    c = prover.prove_resident(tb, pd)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    tb.close()
    pd.close()
    ctx.close()
