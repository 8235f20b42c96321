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


@pytest.mark.parametrize("log_n,width,log_blowup", [(5, 2, 2), (6, 3, 1), (7, 1, 2), (9, 5, 2), (11, 7, 2), (12, 33, 2), (15, 3, 2),
                                                    (17, 2, 2), (18, 3, 1), (19, 1, 1), (20, 2, 1), (21, 1, 1), (22, 1, 1)])
def test_coset_lde_column_kernels_match_tile_kernel(pair, log_n, width, log_blowup):
    """The whole-column LDE kernels (k_ntt_col / k_ntt_top, every group plan 2^5..2^15, one and two passes above stage 15) and the multi-pass tile kernel
    (k_ntt_pass) are independent implementations of the same transform: bit-identical outputs; oracle-checked up to 2^17."""
    ctx, orc = pair
    rng = np.random.default_rng(300 + log_n)
    m = ctx.field.rand(rng, (1 << log_n, width))
    got = ctx.coset_lde(m, log_blowup)
    ctx.set_specialization(1 | 2)  # bit 1: tile kernel only
    try:
        ref = ctx.coset_lde(m, log_blowup)
    finally:
        ctx.set_specialization(1)
    assert np.array_equal(got, ref)
    if log_n <= 17:
        assert np.array_equal(got, orc.coset_lde(m, log_blowup))


@pytest.mark.parametrize("shapes", [[(6, 5)], [(8, 9), (8, 16)], [(9, 3), (7, 20), (7, 1), (4, 11)], [(10, 8), (9, 8), (3, 8)]])
def test_mmcs_commit_mixed_heights(pair, shapes):
    ctx, orc = pair
    rng = np.random.default_rng(5)
    mats = [ctx.field.rand(rng, (1 << lh, w)) for lh, w in shapes]
    assert np.array_equal(ctx.mmcs_commit(mats), orc.mmcs_commit(mats))


@pytest.mark.parametrize("shapes", [[(15, 3), (14, 9), (13, 1), (9, 20), (5, 2)], [(16, 1), (15, 17), (12, 8), (11, 8)],
                                    [(14, 2), (14, 7), (8, 40)]])
def test_mmcs_commit_tall_trees_cross_every_kernel_boundary(pair, shapes):
    """Trees tall enough to use all three Merkle kernels (one-thread-per-node levels above 2^13 nodes, fused k_merkle_stage
    launches below, several stage launches per tree) with rows injected at levels handled by each of them."""
    ctx, orc = pair
    rng = np.random.default_rng(77)
    mats = [ctx.field.rand(rng, (1 << lh, w)) for lh, w in shapes]
    assert np.array_equal(ctx.mmcs_commit(mats), orc.mmcs_commit(mats))


@pytest.mark.parametrize("sizes", [dict(n_const=1, n_public=1, n_alu=1, n_perms=1, n_recompose=1),
                                   dict(n_const=9, n_public=30, n_alu=77, n_perms=0, n_recompose=0),
                                   dict(n_const=40, n_public=3, n_alu=5, n_perms=33, n_recompose=0)])
def test_edge_layers_bit_identical(pair, sizes):
    """Degenerate layers: tables that are almost entirely padding rows (one real operation each), layers without the
    non-primitive tables, and a layer whose widest table is not the tallest."""
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    ctx, orc = pair
    L = wl.synthetic_layer(ctx.field, 5, min_height=16, **sizes)
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    proof = lib.BatchStarkProver(ctx).prove_all_tables(L.traces, pd, L.pubs)
    want = orc.prove(L.insts, L.preps, L.traces, L.pubs)
    assert proof.size == want.size and np.array_equal(proof, want)
    orc.verify(L.insts, pd.preprocessed_commitment, L.pubs, proof)
    pd.close()


def test_isolated_benchmark_entry_points_run(pair):
    """p3r_bench_commit / p3r_bench_fri_round (the isolated sweep of BASELINE.json configs[4]) run and report positive device
    times; argument errors are rejected instead of crashing."""
    ctx, _ = pair
    r = ctx.bench_commit(10, 9, iters=1)
    assert r["lde_ms"] > 0 and r["merkle_ms"] > 0
    f = ctx.bench_fri_round(12, 2, iters=1)
    assert f["fold_ms"] > 0 and f["commit_ms"] > 0
    with pytest.raises(lib.P3RError):
        ctx.bench_fri_round(12, 7, iters=1)   # arity 2^7 is not supported


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


def _fib_system(F, log_n=6):
    fib = importlib.import_module("plonky3-recursion_b200.airs.fibonacci")
    air_mod = importlib.import_module("plonky3-recursion_b200.air")
    t, pubs = fib.trace(F.p, log_n)
    inst = air_mod.build_instance("fib", fib.eval_air, F.p, log_n, 2, 0, 3, air_mod.BusRegistry())
    return [inst], [None], [t], [pubs]


def test_fibonacci_public_values_no_lookups_no_prep(pair):
    ctx, orc = pair
    insts, preps, traces, pubs = _fib_system(ctx.field)
    pd = lib.ProverData.from_airs_and_degrees(ctx, insts, preps)
    assert pd.preprocessed_commitment is None
    proof = lib.BatchStarkProver(ctx).prove_all_tables(traces, pd, pubs)
    assert np.array_equal(proof, orc.prove(insts, preps, traces, pubs))
    orc.verify(insts, None, pubs, proof)
    # an invalid witness still yields a proof, which the verifier rejects (same behaviour as the CPU path)
    bad = traces[0].copy()
    bad[10, 1] = (int(bad[10, 1]) + 1) % ctx.field.p
    bad_proof = lib.BatchStarkProver(ctx).prove_all_tables([bad], pd, pubs)
    with pytest.raises(RuntimeError):
        orc.verify(insts, None, pubs, bad_proof)
    pd.close()


@pytest.mark.parametrize("fri", [
    dict(SMALL_FRI, cap_height=2),
    dict(SMALL_FRI, max_log_arity=1),
    dict(SMALL_FRI, max_log_arity=3, log_final_poly_len=0),
    dict(SMALL_FRI, commit_pow_bits=3, query_pow_bits=0),
    dict(SMALL_FRI, log_blowup=1, num_queries=10),
    dict(SMALL_FRI, log_blowup=3, log_final_poly_len=1),
    dict(lib.DEFAULT_FRI),
])
def test_fri_parameter_variants_bit_identical(fri):
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    ctx = lib.Context("koala-bear", fri)
    orc = make_oracle("koala-bear", fri)
    mh = 256 if fri["log_final_poly_len"] == 5 else 64
    L = wl.synthetic_layer(ctx.field, 4, n_const=10, n_public=20, n_alu=120, n_perms=30, n_recompose=5, min_height=mh)
    fi, fp, ft, fpub = _fib_system(ctx.field, 9)
    insts, preps, traces, pubs = L.insts + fi, L.preps + fp, L.traces + ft, L.pubs + fpub
    pd = lib.ProverData.from_airs_and_degrees(ctx, insts, preps)
    proof = lib.BatchStarkProver(ctx).prove_all_tables(traces, pd, pubs)
    want = orc.prove(insts, preps, traces, pubs)
    assert proof.size == want.size and np.array_equal(proof, want)
    orc.verify(insts, pd.preprocessed_commitment, pubs, proof)
    pd.close()
    ctx.close()


def test_resident_and_pinned_paths_give_the_same_proof(pair):
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    ctx, orc = pair
    L = wl.synthetic_layer(ctx.field, 11, n_const=10, n_public=40, n_alu=300, n_perms=60, n_recompose=5, min_height=32)
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    prover = lib.BatchStarkProver(ctx, pinned_output=True)
    a = prover.prove_all_tables(L.traces, pd, L.pubs)
    b = prover.prove_all_tables(lib.TraceBatch(ctx, L.traces, L.pubs, pinned=True), pd)
    tb = lib.TraceBatch(ctx, L.traces, L.pubs).upload(pd)
    c = prover.prove_resident(tb, pd)
    c2 = prover.prove_resident(tb, pd)  # idempotent: the session must not modify resident traces
    assert np.array_equal(a, b) and np.array_equal(a, c) and np.array_equal(a, c2)
    assert np.array_equal(a, orc.prove(L.insts, L.preps, L.traces, L.pubs))
    tb.close()
    pd.close()


def test_phase_stepped_abi_with_external_transcript(pair):
    """Drive p3r_prove_begin .. p3r_fri_query the way the Rust GpuBatchStarkProver would, with the Fiat-Shamir transcript
    owned by the caller (tests/common.PyChallenger); the assembled proof must equal the oracle's and p3r_prove's."""
    import ctypes as C
    from common import PyChallenger
    abi = importlib.import_module("plonky3-recursion_b200.abi")
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    ctx, orc = pair
    F, l, fri = ctx.field, ctx.lib, ctx.fri
    L = wl.synthetic_layer(F, 21, n_const=10, n_public=30, n_alu=150, n_perms=40, n_recompose=5, min_height=32)
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    tb = lib.TraceBatch(ctx, L.traces, L.pubs)
    sess = C.c_void_p()
    ctx._check(l.p3r_prove_begin(ctx.h, pd.h, tb.tm, tb.pv, C.byref(sess)))
    capw = ctx.cap_words
    mont = lambda xs: np.ascontiguousarray(F.to_monty(np.array(xs, dtype=np.uint32)))
    canon = lambda a: [int(x) for x in F.from_monty(a)]
    ch = PyChallenger(F)
    buf = lambda n: np.zeros(n, dtype=np.uint32)
    main_cap = buf(capw)
    ctx._check(l.p3r_commit_main(sess, abi.as_u32p(main_cap)))
    # a phase called out of order is refused (P3R_ERR_STATE = 4)
    assert l.p3r_commit_quotient(sess, abi.as_u32p(mont([1, 0, 0, 0])), abi.as_u32p(buf(capw))) == 4
    ch.observe_lifted(len(L.insts))
    for s in L.insts:
        for v in (s.log_height, s.log_height, s.main_width, 1 << s.log_quotient_chunks):
            ch.observe_lifted(v)
    ch.observe_many(canon(main_cap))
    for s in L.insts:
        ch.observe_lifted(s.prep_width)
    ch.observe_many(canon(pd.preprocessed_commitment))
    pa, pb = ch.sample_ext(), ch.sample_ext()
    n_perm = sum(1 for s in L.insts if s.lookups)
    perm_cap, terms = buf(capw), buf(4 * n_perm)
    ctx._check(l.p3r_commit_perm(sess, abi.as_u32p(mont(pa)), abi.as_u32p(mont(pb)), abi.as_u32p(perm_cap), abi.as_u32p(terms)))
    ch.observe_many(canon(perm_cap))
    ch.observe_many(canon(terms))
    alpha = ch.sample_ext()
    quot_cap = buf(capw)
    ctx._check(l.p3r_commit_quotient(sess, abi.as_u32p(mont(alpha)), abi.as_u32p(quot_cap)))
    ch.observe_many(canon(quot_cap))
    zeta = ch.sample_ext()
    opened, n_open = buf(1 << 16), C.c_size_t(0)
    ctx._check(l.p3r_open(sess, abi.as_u32p(mont(zeta)), abi.as_u32p(opened), C.c_size_t(opened.size), C.byref(n_open)))
    opened = opened[: n_open.value]
    # observe in round order [main, quotient, preprocessed, permutation]; the blob is per instance
    # [main_local, main_next?, prep_local, prep_next, perm_local, perm_next, quotient chunks]
    segs, cur = [], 0
    for s in L.insts:
        m = 4 * s.main_width * (2 if s.uses_next_row else 1)
        pr = 4 * s.prep_width * 2
        pe = 4 * s.aux_width * 4 * 2
        q = 4 * (4 << s.log_quotient_chunks)
        segs.append((cur, m, pr, pe, q))
        cur += m + pr + pe + q
    assert cur == opened.size
    oc = canon(opened)
    for (o, m, pr, pe, q) in segs:
        ch.observe_many(oc[o:o + m])
    for (o, m, pr, pe, q) in segs:
        ch.observe_many(oc[o + m + pr + pe:o + m + pr + pe + q])
    for (o, m, pr, pe, q) in segs:
        ch.observe_many(oc[o + m:o + m + pr])
    for (o, m, pr, pe, q) in segs:
        ch.observe_many(oc[o + m + pr:o + m + pr + pe])
    alpha_fri = ch.sample_ext()
    n_rounds, arities = C.c_uint32(0), buf(32)
    ctx._check(l.p3r_fri_begin(sess, abi.as_u32p(mont(alpha_fri)), C.byref(n_rounds), abi.as_u32p(arities)))
    R = n_rounds.value
    fri_caps, commit_pow = [], []
    for r in range(R):
        cap = buf(capw)
        ctx._check(l.p3r_fri_commit(sess, r, abi.as_u32p(cap)))
        fri_caps.append(cap)
        ch.observe_many(canon(cap))
        commit_pow.append(0)  # commit_pow_bits == 0: no-op on the transcript
        beta = ch.sample_ext()
        ctx._check(l.p3r_fri_fold(sess, r, abi.as_u32p(mont(beta))))
    final = buf(4 << fri["log_final_poly_len"])
    ctx._check(l.p3r_fri_final_poly(sess, abi.as_u32p(final)))
    ch.observe_many(canon(final))
    for r in range(R):
        ch.observe(int(arities[r]))
    w = ctx.grind(mont(ch.state), mont(ch.inp) if ch.inp else np.zeros(0, dtype=np.uint32), fri["query_pow_bits"])
    w_canon = canon(np.array([w], dtype=np.uint32))[0]
    ch.observe(w_canon)
    assert ch.sample_bits(fri["query_pow_bits"]) == 0
    log_max = max(s.log_height for s in L.insts) + fri["log_blowup"]
    idx = np.array([ch.sample_bits(log_max) for _ in range(fri["num_queries"])], dtype=np.uint32)
    qbuf, nq = buf(1 << 20), C.c_size_t(0)
    ctx._check(l.p3r_fri_query(sess, abi.as_u32p(idx), idx.size, abi.as_u32p(qbuf), C.c_size_t(qbuf.size), C.byref(nq)))
    l.p3r_session_free(sess)
    blob = np.concatenate([
        np.array([0x50335250, len(L.insts), 1, 1, capw] + [s.log_height for s in L.insts], dtype=np.uint32),
        main_cap, perm_cap, quot_cap, terms, opened, np.array([R], dtype=np.uint32), arities[:R], *fri_caps,
        np.array(commit_pow, dtype=np.uint32), final, np.array([w], dtype=np.uint32), qbuf[: nq.value]])
    want = orc.prove(L.insts, L.preps, L.traces, L.pubs)
    assert blob.size == want.size and np.array_equal(blob, want)
    assert np.array_equal(blob, lib.BatchStarkProver(ctx).prove_all_tables(tb, pd))
    pd.close()


@pytest.mark.parametrize("log_n,width", [(18, 3), (20, 2)])
def test_large_lde_properties(pair, log_n, width):
    """Sizes the oracle cannot reach in seconds: check size-independent properties instead. The first n rows of the
    bit-reversed LDE are the evaluations on the coset GENERATOR*H_n (so interpolating them must reproduce the input after
    shifting back), and the LDE is linear."""
    ctx, orc = pair
    F = ctx.field
    rng = np.random.default_rng(log_n)
    a = F.rand(rng, (1 << log_n, width))
    b = F.rand(rng, (1 << log_n, width))
    s = ((a.astype(np.uint64) + b) % F.p).astype(np.uint32)
    la, lb_, ls = ctx.coset_lde(a, 1), ctx.coset_lde(b, 1), ctx.coset_lde(s, 1)
    assert np.array_equal(ls, ((la.astype(np.uint64) + lb_) % F.p).astype(np.uint32))
    # a constant column extends to the same constant; the column x -> x (identity on H_n) extends to GENERATOR*w_N^bitrev(r)
    n, N = 1 << log_n, 1 << (log_n + 1)
    w = F.two_adic_generator(log_n)
    # build the subgroup H_n iteratively in numpy (object-free): powers via cumulative products in chunks
    pw = np.ones(n, dtype=np.uint64)
    step = 1
    base = w
    while step < n:
        pw[step:2 * step] = pw[:step] * np.uint64(base) % np.uint64(F.p)
        base = base * base % F.p
        step *= 2
    # pw currently holds w^(bit-reversed-like enumeration)?  it holds w^i for i in natural order:
    # pw[step + j] = pw[j] * w^(step)  requires base = w^step: base starts at w = w^1 and squares -> w^step. OK.
    ident = np.stack([pw.astype(np.uint32), np.full(n, 7, dtype=np.uint32)], axis=1)
    lde = ctx.coset_lde(ident, 1)
    assert (lde[:, 1] == 7).all()
    wN = F.two_adic_generator(log_n + 1)
    for r in (0, 1, 2, 12345 % N, N - 1):
        rev = int(format(r, f"0{log_n + 1}b")[::-1], 2)
        assert int(lde[rev, 0]) == F.generator * pow(wN, r, F.p) % F.p


def test_specialized_and_interpreted_quotient_agree(pair):
    """The build-time specialised quotient kernels (ALU / Poseidon2 programs) and the bytecode interpreter — in its
    constraint-group schedule and with one thread per row — must produce the same proof; all equal the oracle's."""
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    ctx, orc = pair
    L = wl.synthetic_layer(ctx.field, 31, n_const=10, n_public=40, n_alu=300, n_perms=60, n_recompose=5, min_height=32)
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    prover = lib.BatchStarkProver(ctx)
    ctx.set_specialization(True)
    a = prover.prove_all_tables(L.traces, pd, L.pubs)
    ctx.set_specialization(False)          # interpreter with constraint groups (k_quotient_grouped) for the long programs
    b = prover.prove_all_tables(L.traces, pd, L.pubs)
    ctx.set_specialization(16)             # interpreter, one thread per row (k_quotient)
    c = prover.prove_all_tables(L.traces, pd, L.pubs)
    ctx.set_specialization(True)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    assert np.array_equal(a, orc.prove(L.insts, L.preps, L.traces, L.pubs))
    pd.close()


@pytest.mark.parametrize("lanes", [1, 2])
def test_recompose_with_coefficient_lookups_bit_identical(pair, lanes):
    """`recompose/coeff` table (recompose_air.rs:175-197; 1 + D lookups per lane): one lane runs the generated quotient /
    LogUp kernels, two lanes the bytecode interpreter; both equal the oracle, with and without specialisation."""
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    ctx, orc = pair
    L = wl.synthetic_layer(ctx.field, 41, n_const=12, n_public=50, n_alu=300, n_perms=40, n_recompose=37, min_height=32,
                           recompose_coeff=True, recompose_lanes=lanes)
    assert L.insts[-1].name == "recompose/coeff" and L.insts[-1].aux_width == 1 + 5 * lanes
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    prover = lib.BatchStarkProver(ctx)
    want = orc.prove(L.insts, L.preps, L.traces, L.pubs)
    a = prover.prove_all_tables(L.traces, pd, L.pubs)
    ctx.set_specialization(False)
    b = prover.prove_all_tables(L.traces, pd, L.pubs)
    ctx.set_specialization(True)
    assert np.array_equal(a, want) and np.array_equal(b, want)
    orc.verify(L.insts, pd.preprocessed_commitment, L.pubs, a)
    pd.close()


@pytest.mark.parametrize("n", [1000, 40])
def test_base_layer_fibonacci_circuit_bit_identical(pair, n):
    """BASELINE configs[0]: the extension-degree-1 base circuit of recursive_fibonacci (ALU 1024 x 7, Const / Public 256 x 1,
    WitnessChecks tuples of width 2) — interpreter path for every table, equal to the oracle bit for bit."""
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    ctx, orc = pair
    L = wl.base_layer_fibonacci(ctx.field, n, min_height=256 if n == 1000 else 16)
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    proof = lib.BatchStarkProver(ctx).prove_all_tables(L.traces, pd, L.pubs)
    assert np.array_equal(pd.preprocessed_commitment, orc.prep_commit(L.insts, L.preps))
    assert np.array_equal(proof, orc.prove(L.insts, L.preps, L.traces, L.pubs))
    orc.verify(L.insts, pd.preprocessed_commitment, L.pubs, proof)
    pd.close()


def test_gpu_alu_table_fill_matches_reference_builder(pair):
    """The device-generated ALU table (schedule slots + operand values -> 80 columns incl. packed-Horner intermediates,
    (a_t, c_t) operands and b^2) equals the host restatement of AluAir::trace_to_matrix bit for bit; proofs from operation
    lists (ALU + Poseidon2 both generated on the device) equal proofs from uploaded matrices and the oracle's."""
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    ctx, orc = pair
    for seed, n_alu in ((51, 300), (52, 1), (53, 97)):
        L = wl.synthetic_layer(ctx.field, seed, n_const=10, n_public=40, n_alu=n_alu, n_perms=20, n_recompose=5, min_height=32)
        (idx, _), = L.alu_ops.items()
        assert (L.alu_ops[idx].slot_kind >= 2).any() or n_alu == 1
        pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
        tb = lib.TraceBatch(ctx, L.traces, L.pubs, p2_ops=L.p2_ops, alu_ops=L.alu_ops).upload(pd)
        got = tb.download(pd, idx)
        assert got.shape == L.traces[idx].shape and np.array_equal(got, L.traces[idx])
        prover = lib.BatchStarkProver(ctx)
        from_ops = prover.prove_all_tables(lib.TraceBatch(ctx, L.traces, L.pubs, p2_ops=L.p2_ops, alu_ops=L.alu_ops), pd)
        pinned_ops = prover.prove_all_tables(lib.TraceBatch(ctx, L.traces, L.pubs, pinned=True, p2_ops=L.p2_ops, alu_ops=L.alu_ops), pd)
        assert np.array_equal(from_ops, prover.prove_all_tables(L.traces, pd, L.pubs))
        assert np.array_equal(from_ops, pinned_ops)
        assert np.array_equal(from_ops, prover.prove_resident(tb, pd))
        assert np.array_equal(from_ops, orc.prove(L.insts, L.preps, L.traces, L.pubs))
        tb.close()
        pd.close()


def test_device_and_host_fri_transcripts_agree(pair):
    """p3r_prove samples the FRI commit-phase betas on the device (k_fri_round_transcript) and replays them on the host
    challenger; with bit 2 of p3r_set_specialization the host samples them round by round. Same proof either way, equal to the
    oracle's; every code-path switch combined (interpreter + tile-kernel LDE + host transcript) also gives the same bytes."""
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    ctx, orc = pair
    L = wl.synthetic_layer(ctx.field, 33, n_const=12, n_public=50, n_alu=150, n_perms=40, n_recompose=6, min_height=32)
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    prover = lib.BatchStarkProver(ctx)
    want = orc.prove(L.insts, L.preps, L.traces, L.pubs)
    try:
        for flags in (1, 1 | 4, 0 | 2 | 4, 1 | 2):
            ctx.set_specialization(flags)
            assert np.array_equal(prover.prove_all_tables(L.traces, pd, L.pubs), want), flags
    finally:
        ctx.set_specialization(1)
    pd.close()


def test_gpu_poseidon2_table_fill_matches_reference_builder(pair):
    """K3: the device-generated Poseidon2 table equals the host restatement of generate_trace_rows bit for bit (padding rows,
    Merkle index accumulator, S-box registers for BabyBear), and proofs from ops == proofs from the uploaded matrix."""
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    ctx, orc = pair
    for seed, n_perms in ((41, 70), (42, 128), (43, 1)):
        L = wl.synthetic_layer(ctx.field, seed, n_const=10, n_public=40, n_alu=200, n_perms=n_perms, n_recompose=5, min_height=32)
        (idx, ops), = L.p2_ops.items()
        pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
        tb = lib.TraceBatch(ctx, L.traces, L.pubs, p2_ops=L.p2_ops).upload(pd)
        got = tb.download(pd, idx)
        assert got.shape == L.traces[idx].shape and np.array_equal(got, L.traces[idx])
        prover = lib.BatchStarkProver(ctx)
        from_ops = prover.prove_all_tables(lib.TraceBatch(ctx, L.traces, L.pubs, p2_ops=L.p2_ops), pd)
        from_matrix = prover.prove_all_tables(L.traces, pd, L.pubs)
        assert np.array_equal(from_ops, from_matrix)
        assert np.array_equal(from_ops, prover.prove_resident(tb, pd))
        assert np.array_equal(from_ops, orc.prove(L.insts, L.preps, L.traces, L.pubs))
        tb.close()
        pd.close()


def test_full_size_layer_is_accepted_by_the_oracle_verifier():
    """BASELINE.json's headline configuration (the KoalaBear steady-state layer `bench.py` times: ALU 2^15 x 80, Poseidon2
    2^14 x 166, Public 2^16 x 4, production FRI parameters): the oracle VERIFIER accepts the GPU proof; the oracle PROVER's
    proof of the same inputs is word for word the GPU's (through the oracle's CPU-arm routes, which
    tests/test_oracle_fast_paths.py ties to its plain ones — the plain prover alone needs ~5 s per proof at this size);
    proving is deterministic; matrices, pinned matrices, device-resident traces and operation lists (tables generated on the
    device) give the same bytes; the specialised and interpreted paths agree; a flipped word anywhere is rejected."""
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    ctx = lib.Context("koala-bear", lib.DEFAULT_FRI)
    orc = make_oracle("koala-bear", lib.DEFAULT_FRI)
    L = wl.synthetic_layer(ctx.field, 1, n_const=1500, n_public=43000, n_alu=60000, n_perms=12000, n_recompose=4000,
                           min_height=256)
    assert [s[1] for s in L.shapes] == [2048, 65536, 32768, 16384, 4096]
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    prover = lib.BatchStarkProver(ctx)
    proof = prover.prove_all_tables(L.traces, pd, L.pubs)
    orc.verify(L.insts, pd.preprocessed_commitment, L.pubs, proof)
    if orc.set_fast_paths(True):
        assert np.array_equal(proof, orc.prove(L.insts, L.preps, L.traces, L.pubs))     # headline size, bit for bit
    assert np.array_equal(proof, prover.prove_all_tables(L.traces, pd, L.pubs))
    tb = lib.TraceBatch(ctx, L.traces, L.pubs).upload(pd)
    assert np.array_equal(proof, prover.prove_resident(tb, pd))
    tb.close()
    ops = lib.TraceBatch(ctx, L.traces, L.pubs, pinned=True, p2_ops=L.p2_ops, alu_ops=L.alu_ops)
    assert np.array_equal(proof, prover.prove_all_tables(ops, pd))
    ctx.set_specialization(False)
    assert np.array_equal(proof, prover.prove_all_tables(L.traces, pd, L.pubs))
    ctx.set_specialization(True)
    rng = np.random.default_rng(0)
    for pos in [3, 20, proof.size // 3, proof.size // 2, proof.size - 9] + [int(x) for x in rng.integers(0, proof.size, 6)]:
        bad = proof.copy()
        bad[pos] = (int(bad[pos]) + 1) % ctx.field.p
        with pytest.raises(RuntimeError):
            orc.verify(L.insts, pd.preprocessed_commitment, L.pubs, bad)
    pd.close()
    ctx.close()


@pytest.mark.parametrize("public_lanes,alu_lanes,horner_k", [(2, 2, 4), (2, 4, 3), (4, 3, 2)])
def test_table_packing_variants_bit_identical(pair, public_lanes, alu_lanes, horner_k):
    """Other `TablePacking` shapes (public lanes, ALU lanes, packed-Horner depth): interpreter kernels for every table."""
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    ctx, orc = pair
    L = wl.synthetic_layer(ctx.field, 13, n_const=10, n_public=33, n_alu=260, n_perms=20, n_recompose=5, min_height=16,
                           public_lanes=public_lanes, alu_lanes=alu_lanes, horner_k=horner_k)
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    proof = lib.BatchStarkProver(ctx).prove_all_tables(L.traces, pd, L.pubs)
    assert np.array_equal(proof, orc.prove(L.insts, L.preps, L.traces, L.pubs))
    orc.verify(L.insts, pd.preprocessed_commitment, L.pubs, proof)
    pd.close()


AGG_FRI = dict(log_blowup=2, log_final_poly_len=6, max_log_arity=2, num_queries=5, commit_pow_bits=0, query_pow_bits=5, cap_height=0)


@pytest.mark.parametrize("field_name", FIELDS)
@pytest.mark.parametrize("public_lanes,alu_lanes", [(1, 2), (1, 1), (2, 2)])
def test_reference_example_packings_bit_identical(field_name, public_lanes, alu_lanes):
    """The packings BASELINE configs[2]/[3] use: `TablePacking::new(1, 2)` (recursion/examples/recursive_keccak.rs:562),
    `new(1, 1)` for base proofs and `new(2, 2)` for aggregation level 1 (recursive_aggregation.rs:632,666) — default packed
    Horner depth 2 (batch_stark_prover/packing.rs:28-30) — with the aggregation example's `log_final_poly_len` 6
    (recursive_aggregation.rs:83), hence min_trace_height 2^(6+2+1) = 512 (packing.rs:100-106)."""
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    ctx = lib.Context(field_name, AGG_FRI)
    orc = make_oracle(field_name, AGG_FRI)
    L = wl.synthetic_layer(ctx.field, 17, n_const=12, n_public=40, n_alu=700, n_perms=30, n_recompose=6, min_height=512,
                           public_lanes=public_lanes, alu_lanes=alu_lanes, horner_k=2)
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    proof = lib.BatchStarkProver(ctx).prove_all_tables(L.traces, pd, L.pubs)
    assert np.array_equal(proof, orc.prove(L.insts, L.preps, L.traces, L.pubs))
    orc.verify(L.insts, pd.preprocessed_commitment, L.pubs, proof)
    pd.close()
    ctx.close()


def test_contexts_with_different_poseidon2_constants_are_refused():
    """The thread-per-permutation kernels read the round constants from one `__constant__` slot per (device, field): a second
    live context of the same field with other constants must fail loudly (P3R_ERR_UNSUPPORTED) instead of corrupting the
    first one's hashes; the same constants share the slot; after the first context is gone the other constants load."""
    p2mod = importlib.import_module("plonky3-recursion_b200.poseidon2_params")
    F = field_mod.get_field("koala-bear")
    a = lib.Context("koala-bear", SMALL_FRI)
    orc = make_oracle("koala-bear", SMALL_FRI)
    st = F.rand(np.random.default_rng(5), (33, 16))
    want = orc.poseidon2_permute(st)
    other = p2mod.Poseidon2Params(F.field_id)
    other.external_rc = other.external_rc.copy()
    other.external_rc[0] = (int(other.external_rc[0]) + 1) % F.p
    with pytest.raises(lib.P3RError) as ei:
        lib.Context("koala-bear", SMALL_FRI, poseidon2=other)
    assert ei.value.code == 5
    b = lib.Context("koala-bear", SMALL_FRI)              # same constants: shares the slot
    assert np.array_equal(a.poseidon2_permute(st), want) and np.array_equal(b.poseidon2_permute(st), want)
    a.close()
    b.close()
    c = lib.Context("koala-bear", SMALL_FRI, poseidon2=other)   # nobody else alive: the other constants load
    got = c.poseidon2_permute(st)
    assert not np.array_equal(got, want)
    assert np.array_equal(got, other.permute(st))
    c.close()
    d = lib.Context("koala-bear", SMALL_FRI)
    assert np.array_equal(d.poseidon2_permute(st), want)
    d.close()


def test_missing_public_values_are_an_invalid_argument():
    """n_public > 0 with a NULL public-values pointer is P3R_ERR_INVALID_ARG at the C ABI (not a host crash)."""
    import ctypes as C
    abi = importlib.import_module("plonky3-recursion_b200.abi")
    fib = importlib.import_module("plonky3-recursion_b200.airs.fibonacci")
    air_mod = importlib.import_module("plonky3-recursion_b200.air")
    ctx = lib.Context("koala-bear", SMALL_FRI)
    F = ctx.field
    inst = air_mod.build_instance("fib", fib.eval_air, F.p, 5, 2, 0, 3, air_mod.BusRegistry())
    pd = lib.ProverData.from_airs_and_degrees(ctx, [inst], [None])
    trace, _ = fib.trace(F.p, 5)
    with pytest.raises(ValueError):
        lib.BatchStarkProver(ctx).prove_all_tables([trace], pd, [None])      # the Python mirror checks lengths first
    m = abi.Marshal(F)
    tm = m.matrices([trace])
    pv = (abi.u32p * 1)()                                                      # NULL entry
    buf = np.zeros(1 << 16, dtype=np.uint32)
    n = C.c_size_t(0)
    rc = ctx.lib.p3r_prove(ctx.h, pd.h, tm, pv, abi.as_u32p(buf), C.c_size_t(buf.size), C.byref(n))
    assert rc == 1 and b"public values" in ctx.lib.p3r_last_error(ctx.h)
    rc = ctx.lib.p3r_prove(ctx.h, pd.h, tm, None, abi.as_u32p(buf), C.c_size_t(buf.size), C.byref(n))
    assert rc == 1
    pd.close()
    ctx.close()


def test_work_queue_row_hashing_matches_one_cta_per_rows(pair):
    """k_hash_rows (one CTA per 64 rows, the product path) and k_hash_rows_queue (32-row work items taken from an atomic
    counter, longest sponges first) are two schedules of the same sponge: identical mixed-height commitments, both equal to
    the oracle's."""
    ctx, orc = pair
    rng = np.random.default_rng(77)
    mats = [ctx.field.rand(rng, (1 << lh, w)) for lh, w in [(12, 3), (11, 70), (10, 170), (12, 9), (7, 5), (3, 33)]]
    want = orc.mmcs_commit(mats)
    got_cta = ctx.mmcs_commit(mats)
    ctx.set_specialization(1 | 8)
    try:
        got_queue = ctx.mmcs_commit(mats)
    finally:
        ctx.set_specialization(1)
    assert np.array_equal(got_queue, want) and np.array_equal(got_cta, want)



@pytest.mark.gpu
def test_full_size_babybear_layer_matches_the_oracle_prover():
    """The headline layer over BabyBear (Poseidon2 table 2^14 x 300, S-box degree 7): GPU proof == oracle prover's proof,
    all words, and the oracle verifier accepts it."""
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    ctx = lib.Context("baby-bear", lib.DEFAULT_FRI)
    orc = make_oracle("baby-bear", lib.DEFAULT_FRI)
    if not orc.set_fast_paths(True):
        ctx.close()
        pytest.skip("the plain oracle prover needs several seconds per proof at this size")
    L = wl.synthetic_layer(ctx.field, 1, n_const=1500, n_public=43000, n_alu=60000, n_perms=12000, n_recompose=4000,
                           min_height=256)
    assert [s[1:3] for s in L.shapes] == [(2048, 4), (65536, 4), (32768, 80), (16384, 300), (4096, 4)]
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    proof = lib.BatchStarkProver(ctx).prove_all_tables(L.traces, pd, L.pubs)
    orc.verify(L.insts, pd.preprocessed_commitment, L.pubs, proof)
    assert np.array_equal(proof, orc.prove(L.insts, L.preps, L.traces, L.pubs))
    pd.close()
    ctx.close()
