"""Pins of the CPU oracle: every known-answer / structural fact available for this path.

The reference tree carries no golden vectors for commitments or proofs ("parity unpinned" at the byte level); what can be
pinned is pinned here: field parameters, the Poseidon2 round constants against the p3 tables as recalled, the oracle's
primitives against independent numpy / textbook computations, the challenger semantics of recursion/src/challenger/circuit.rs."""
import importlib

import numpy as np
import pytest

from common import SMALL_FRI, field_mod, make_oracle, p2mod

FIELDS = ["koala-bear", "baby-bear"]


def test_field_parameters():
    kb, bb = field_mod.get_field("koala-bear"), field_mod.get_field("baby-bear")
    # moduli asserted by the reference's own tests (circuit-prover/src/batch_stark_prover/tests.rs:704-705,731-732)
    assert kb.p == 0x7F000001 == 2130706433 and bb.p == 0x78000001 == 2013265921
    for F in (kb, bb):
        # generator has full order p-1: g^((p-1)/q) != 1 for every prime q | p-1
        n = F.p - 1
        primes = [q for q in (2, 3, 5, 7, 127) if n % q == 0]
        m = n
        for q in primes:
            while m % q == 0:
                m //= q
        assert m == 1, "unexpected prime factor of p-1"
        assert all(pow(F.generator, n // q, F.p) != 1 for q in primes)
        assert n % (1 << F.two_adicity) == 0 and (n >> F.two_adicity) % 2 == 1
        # two_adic_generator(k) = g^((p-1)/2^k): order exactly 2^k
        w = F.two_adic_generator(F.two_adicity)
        assert pow(w, 1 << F.two_adicity, F.p) == 1 and pow(w, 1 << (F.two_adicity - 1), F.p) == F.p - 1
        # W is a quartic non-residue => x^4 - W irreducible (binomial extension, circuit-prover/src/field_params.rs:34-41)
        assert pow(F.w, (F.p - 1) // 2, F.p) == F.p - 1
    assert pow(31, 15, bb.p) == 0x1A427A41 and pow(3, 127, kb.p) == 0x6AC49F88  # SURVEY.md §7 H2 cross-check
    # Montgomery constants
    assert pow(kb.p, -1, 1 << 32) == 0x81000001 and (1 << 32) % kb.p == 0x01FFFFFE and (1 << 64) % kb.p == 0x17F7EFE4
    assert pow(bb.p, -1, 1 << 32) == 0x88000001 and (1 << 32) % bb.p == 0x0FFFFFFE and (1 << 64) % bb.p == 0x45DDDDE3


def test_poseidon2_round_constants_match_p3_tables():
    """Grain-LFSR regeneration vs the p3 constant tables as recalled (BabyBear: horizen-labs derived RC16 table)."""
    bb = p2mod.Poseidon2Params(field_mod.BABYBEAR)
    assert [int(x) for x in bb.external_rc[:16]] == [
        0x69CBB6AF, 0x46AD93F9, 0x60A00F4E, 0x6B1297CD, 0x23189AFE, 0x732E7BEF, 0x72C246DE, 0x2C941900,
        0x0557EEDE, 0x1580496F, 0x3A3EA77B, 0x54F3F271, 0x0F49B029, 0x47872FE1, 0x221E2E36, 0x1AB7202E]
    assert [int(x) for x in bb.internal_rc] == [
        0x5A8053C0, 0x693BE639, 0x3858867D, 0x19334F6B, 0x128F0FD8, 0x4E2B1CCB, 0x61210CE0, 0x3C318939,
        0x0B5B2F22, 0x2EDB11D5, 0x213EFFDF, 0x0CAC4606, 0x241AF16D]
    kb = p2mod.Poseidon2Params(field_mod.KOALABEAR)
    assert [int(x) for x in kb.external_rc[:8]] == [
        0x7EE56A48, 0x11367045, 0x12E41941, 0x7EBBC12B, 0x1970B7D5, 0x662B60E8, 0x3E4990C6, 0x679F91F5]
    assert (kb.rounds_f, kb.rounds_p, kb.sbox_degree) == (8, 20, 3)  # circuit/src/ops/poseidon2_perm/config.rs:114-122
    assert (bb.rounds_f, bb.rounds_p, bb.sbox_degree) == (8, 13, 7)  # :67-75
    for prm in (kb, bb):
        p = prm.field.p
        # internal diagonal entries as documented in SURVEY.md §8c: -2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, then +-1/2^k
        d = [int(x) for x in prm.internal_diag]
        assert d[:3] == [p - 2, 1, 2] and d[3] * 2 % p == 1 and d[4:6] == [3, 4] and (d[6] * 2 + 1) % p == 0
        assert d[7:9] == [p - 3, p - 4] and d[9] * 256 % p == 1
        assert len(set(d)) == 16 and np.gcd(prm.sbox_degree, p - 1) == 1  # S-box is a permutation


@pytest.mark.parametrize("field", FIELDS)
def test_oracle_permutation_equals_numpy_reference(field):
    orc = make_oracle(field)
    prm = p2mod.Poseidon2Params(orc.field.field_id)
    rng = np.random.default_rng(0)
    st = orc.field.rand(rng, (33, 16))
    st[0] = 0
    got = orc.poseidon2_permute(st)
    assert np.array_equal(got, prm.permute(st))
    assert len({tuple(r) for r in got}) == 33  # injective on the sample
    # the permutation is not linear: perm(a+b) != perm(a)+perm(b)
    a, b = st[1].astype(np.uint64), st[2].astype(np.uint64)
    s = ((a + b) % orc.field.p).astype(np.uint32)
    lin = ((got[1].astype(np.uint64) + got[2]) % orc.field.p).astype(np.uint32)
    assert not np.array_equal(orc.poseidon2_permute(s[None])[0], lin)


@pytest.mark.parametrize("field", FIELDS)
def test_oracle_lde_is_polynomial_evaluation(field):
    """coset_lde output row bitrev(r) == P(GENERATOR * w_N^r) with P interpolating the column over H_n."""
    orc = make_oracle(field)
    F = orc.field
    p = F.p
    rng = np.random.default_rng(4)
    log_n, lb = 4, 2
    n, N = 1 << log_n, 1 << (log_n + lb)
    col = [int(x) for x in F.rand(rng, n)]
    w = F.two_adic_generator(log_n)
    # textbook O(n^2) interpolation: coeff_k = (1/n) sum_i v_i w^{-ik}
    coeffs = [sum(col[i] * pow(w, (-i * k) % n, p) for i in range(n)) * F.inv(n) % p for k in range(n)]
    wN = F.two_adic_generator(log_n + lb)
    out = orc.coset_lde(np.array(col, dtype=np.uint32).reshape(n, 1), lb)[:, 0]
    for r in range(N):
        x = F.generator * pow(wN, r, p) % p
        want = sum(c * pow(x, k, p) for k, c in enumerate(coeffs)) % p
        rev = int(format(r, f"0{log_n + lb}b")[::-1], 2)
        assert int(out[rev]) == want
    # linearity of the LDE
    a, b = F.rand(rng, (n, 3)), F.rand(rng, (n, 3))
    s = ((a.astype(np.uint64) + b) % p).astype(np.uint32)
    la, lb_, ls = orc.coset_lde(a, lb), orc.coset_lde(b, lb), orc.coset_lde(s, lb)
    assert np.array_equal(ls, ((la.astype(np.uint64) + lb_) % p).astype(np.uint32))


@pytest.mark.parametrize("field", FIELDS)
def test_mmcs_open_verify_mixed_heights(field):
    """Every leaf of a mixed-height tree opens and verifies (injection semantics of SURVEY.md A8); edge shapes included."""
    orc = make_oracle(field)
    rng = np.random.default_rng(9)
    F = orc.field
    for shapes in ([(3, 5)], [(4, 9), (4, 1)], [(5, 3), (3, 20), (3, 1), (1, 11)], [(4, 8), (3, 8), (0, 8)]):
        mats = [F.rand(rng, (1 << lh, w)) for lh, w in shapes]
        for idx in range(1 << shapes[0][0]):
            orc.mmcs_open_verify(mats, idx)
    # width exactly a multiple of the sponge rate, width 1, and changing any entry changes the root
    m = F.rand(rng, (8, 16))
    root = orc.mmcs_commit([m])
    m2 = m.copy()
    m2[5, 15] = (int(m2[5, 15]) + 1) % F.p
    assert not np.array_equal(root, orc.mmcs_commit([m2]))
    # order of equal-height matrices matters (leaf = concatenation in commit order)
    a, b = F.rand(rng, (8, 3)), F.rand(rng, (8, 3))
    assert not np.array_equal(orc.mmcs_commit([a, b]), orc.mmcs_commit([b, a]))


@pytest.mark.parametrize("field", FIELDS)
def test_duplex_challenger_semantics(field):
    """recursion/src/challenger/circuit.rs:97-156,337-430: overwrite absorb, zero-fill, length tag into state[8], samples popped
    from the back of the rate, sampling after an observe re-duplexes."""
    orc = make_oracle(field)
    F = orc.field
    prm = p2mod.Poseidon2Params(F.field_id)
    rng = np.random.default_rng(2)
    xs = F.rand(rng, 11)
    mon = F.to_monty(xs)
    # observe 3, sample 2 : one duplex with n=3
    out = F.from_monty(orc.challenger_script([(0, 0)] * 3 + [(1, 0)] * 2, mon[:3]))
    st = np.zeros(16, dtype=np.uint64)
    st[:3] = xs[:3]
    st[8] = 3
    perm = prm.permute(st[None])[0]
    assert list(out) == [perm[7], perm[6]]
    # observe 8 (auto duplex, tag 8), then 2 more, sample: second duplex overwrites [0,2), zeroes [2,8), tag += 2
    out = F.from_monty(orc.challenger_script([(0, 0)] * 10 + [(1, 0)], mon[:10]))
    st = np.zeros(16, dtype=np.uint64)
    st[:8] = xs[:8]
    st[8] = 8
    s1 = prm.permute(st[None])[0].astype(np.uint64)
    s1[:2] = xs[8:10]
    s1[2:8] = 0
    s1[8] = (s1[8] + 2) % F.p
    assert int(out[0]) == int(prm.permute(s1[None])[0][7])
    # sample_bits = low bits of the canonical sample; 9 samples force a squeeze-only duplex (state permuted untouched)
    ops = [(0, 0)] + [(1, 0)] * 9
    out = F.from_monty(orc.challenger_script(ops, mon[:1]))
    st = np.zeros(16, dtype=np.uint64)
    st[0] = xs[0]
    st[8] = 1
    p1 = prm.permute(st[None])[0]
    p2 = prm.permute(p1[None].astype(np.uint64))[0]
    assert list(out[:8]) == list(p1[:8][::-1]) and int(out[8]) == int(p2[7])
    bits = orc.challenger_script([(0, 0), (2, 5)], mon[:1])
    assert int(bits[0]) == int(p1[7]) & 31


def test_grind_witness_is_minimal_and_valid():
    orc = make_oracle("koala-bear")
    F = orc.field
    prm = p2mod.Poseidon2Params(F.field_id)
    rng = np.random.default_rng(5)
    state = F.rand(rng, 16)
    pending = F.rand(rng, 2)
    bits = 6
    w = int(F.from_monty(np.array([orc.grind(F.to_monty(state), F.to_monty(pending), bits)]))[0])

    def sample_after(wit):
        st = state.astype(np.uint64).copy()
        st[:2] = pending
        st[2] = wit
        st[3:8] = 0
        st[8] = (st[8] + 3) % F.p
        return int(prm.permute(st[None])[0][7])

    assert sample_after(w) & ((1 << bits) - 1) == 0
    assert all(sample_after(v) & ((1 << bits) - 1) != 0 for v in range(w))
