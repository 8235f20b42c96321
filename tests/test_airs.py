"""AIR restatements vs the reference's own pins: column-count goldens (circuit-prover/src/air/shape_golden.rs:33-68),
constraint degree <= 3 (test-utils/src/lib.rs:10,101), constraint satisfaction on hand-built traces and rejection of
tampered ones (the `check_air_satisfies` / `assert_air_rejects` pattern of test-utils/src/lib.rs:155-300)."""
import importlib

import numpy as np
import pytest

from common import SMALL_FRI, air_mod, field_mod, make_oracle, p2mod, ws

alu = importlib.import_module("plonky3-recursion_b200.airs.alu")
p2air = importlib.import_module("plonky3-recursion_b200.airs.poseidon2")
sym = importlib.import_module("plonky3-recursion_b200.symbolic")
wl = importlib.import_module("plonky3-recursion_b200.workload")

MAX_TEST_CONSTRAINT_DEGREE = 3


def test_shape_goldens():
    # shape_golden.rs: const D1 (1,2) / D4 (4,2) / D5 (5,2); public D1x1 (1,2), D4x1 (4,2), D4x2 (8,4)
    assert ws.widths(1, 1) == (1, 2) and ws.widths(4, 1) == (4, 2) and ws.widths(5, 1) == (5, 2)
    assert ws.widths(4, 2) == (8, 4)
    # ALU: D1 lane1 (7,20), lane2 (11,33); D2 (14,20); D4 lane1 (28,20), lane2 (44,33); D5 (35,20)
    assert alu.widths(1, 1) == (7, 20) and alu.widths(1, 2) == (11, 33)
    assert alu.widths(2, 1) == (14, 20) and alu.widths(4, 1) == (28, 20) and alu.widths(4, 2) == (44, 33)
    assert alu.widths(5, 1) == (35, 20)
    # recursion-layer packing: 3 lanes, D=4, k=4 -> 80 main / 60 preprocessed
    assert alu.widths(4, 3, 4) == (80, 60)
    # Poseidon2 circuit table: 164/298 permutation columns + mmcs_bit + mmcs_index_sum; 24 preprocessed
    assert p2air.widths(p2mod.Poseidon2Params(field_mod.KOALABEAR)) == (166, 24)
    assert p2air.widths(p2mod.Poseidon2Params(field_mod.BABYBEAR)) == (300, 24)


@pytest.mark.parametrize("field", ["koala-bear", "baby-bear"])
def test_constraint_degrees_and_lookup_packing(field):
    F = field_mod.get_field(field)
    buses = air_mod.BusRegistry()
    prm = p2mod.Poseidon2Params(F.field_id)
    aw, apw = alu.widths(4, 3, 4)
    insts = {
        "const": air_mod.build_instance("const", ws.make_eval(4, 1), F.p, 8, 4, 2, 0, buses),
        "alu": air_mod.build_instance("alu", alu.make_eval(4, 3, 4, F.w), F.p, 8, aw, apw, 0, buses),
        "p2": air_mod.build_instance("p2", p2air.make_eval(prm), F.p, 8, *p2air.widths(prm), 0, buses),
    }
    # tables without local constraints need one quotient chunk, degree-3 tables two
    assert insts["const"].log_quotient_chunks == 0
    assert insts["alu"].log_quotient_chunks == 1 and insts["p2"].log_quotient_chunks == 1
    # interactions: ALU lanes*4 + 2(k-1) = 18 (alu_air.rs:60), Poseidon2 4 in + 2 out + 1 mmcs = 7 (air.rs:1798-1894)
    assert len(insts["alu"].interactions) == 18 and len(insts["p2"].interactions) == 7 and len(insts["const"].interactions) == 1
    assert all(l[0] == 0 for s in insts.values() for l in s.lookups)  # one global bus: WitnessChecks
    assert insts["const"].uses_next_row is False and insts["alu"].uses_next_row and insts["p2"].uses_next_row
    for name, ev, mw, pw in (("alu", alu.make_eval(4, 3, 4, F.w), aw, apw), ("p2", p2air.make_eval(prm), *p2air.widths(prm))):
        b = sym.AirBuilder(F.p, mw, pw, 0)
        ev(b)
        assert sym.max_constraint_degree(b) <= MAX_TEST_CONSTRAINT_DEGREE, name
    # Poseidon2Air constraint count: 8 full rounds x 16 + partial rounds (+ registers) + 34 circuit-level
    n_p2 = 8 * 16 + prm.rounds_p + (8 * 16 + prm.rounds_p) * p2air.sbox_registers(prm) + 1 + 16 + 8 + 8 + 1
    b = sym.AirBuilder(F.p, *p2air.widths(prm), 0)
    p2air.make_eval(prm)(b)
    assert len(b.base_constraints) == n_p2


def _layer(F, seed=3, **kw):
    args = dict(n_const=12, n_public=20, n_alu=150, n_perms=40, n_recompose=6, min_height=32)
    args.update(kw)
    return wl.synthetic_layer(F, seed, **args)


@pytest.mark.parametrize("field", ["koala-bear", "baby-bear"])
def test_synthetic_layer_satisfies_every_air(field):
    orc = make_oracle(field)
    L = _layer(orc.field)
    assert [s[0] for s in L.shapes] == ["const", "public", "alu", "poseidon2", "recompose"]
    for s, pm, tr in zip(L.insts, L.preps, L.traces):
        assert orc.check_constraints(s, pm, tr, None) is None, s.name


def test_alu_rejects_wrong_results_for_every_op_kind():
    orc = make_oracle("koala-bear")
    F = orc.field
    L = _layer(F, seed=5, n_perms=0, n_recompose=0)
    inst, prep, trace = L.insts[2], L.preps[2], L.traces[2]
    assert orc.check_constraints(inst, prep, trace, None) is None
    seen = set()
    for row in range(trace.shape[0]):
        for lane in range(3):
            p = prep[row, lane * 13:(lane + 1) * 13]
            if p[alu.MULT_A] == 0:
                continue
            kind = ("add" if p[alu.SEL_ADD] else "bool" if p[alu.SEL_BOOL] else "muladd" if p[alu.SEL_MULADD] else
                    "horner" if p[alu.SEL_HORNER] else "mul")
            if kind in seen:
                continue
            seen.add(kind)
            bad = trace.copy()
            col = lane * 16 + (0 if kind == "bool" else 12)  # break `out` (or `a` for the boolean check)
            bad[row, col] = (int(bad[row, col]) + 2) % F.p
            res = orc.check_constraints(inst, prep, bad, None)
            assert res is not None and res[0] in (row, (row - 1) % trace.shape[0]), (kind, res)
    assert seen == {"add", "mul", "bool", "muladd", "horner"}
    # packed Horner rows exist and are covered (sel_k for k in 2..4 somewhere in the preprocessed trace)
    extra = 3 * 13
    assert prep[:, extra:extra + 3].sum() > 0


def test_poseidon2_table_semantics():
    """Sponge chaining, Merkle left/right placement, MMCS index accumulator (poseidon2-circuit-air/src/air.rs:1898-2529)."""
    orc = make_oracle("koala-bear")
    F = orc.field
    L = _layer(F, seed=8, n_perms=80)
    inst, prep, trace = L.insts[3], L.preps[3], L.traces[3]
    prm = p2mod.Poseidon2Params(F.field_id)
    lay = p2air.Layout(prm)
    assert orc.check_constraints(inst, prep, trace, None) is None
    n_ops = int((prep[:, :].any(axis=1)).sum())
    # every row (padding included) carries a real permutation: outputs == Poseidon2(inputs)
    assert np.array_equal(trace[:, lay.out_base:lay.out_base + 16], prm.permute(trace[:, :16]))
    # first padding row restarts the chain, later padding rows are all-zero preprocessed (air.rs:613-649)
    pad0 = np.nonzero(~prep.any(axis=1))[0]
    first_pad = int(np.nonzero(prep[:, p2air.NEW_START])[0].max())
    assert (prep[first_pad, :p2air.NEW_START] == 0).all() and first_pad + 1 in pad0 or first_pad == trace.shape[0] - 1
    merkle_rows = np.nonzero(prep[:, p2air.MERKLE_PATH])[0]
    assert merkle_rows.size > 0
    r = int(merkle_rows[len(merkle_rows) // 2])
    # flipping the direction bit breaks the left/right chaining constraint
    bad = trace.copy()
    bad[r, lay.mmcs_bit] ^= 1
    assert orc.check_constraints(inst, prep, bad, None) is not None
    # a wrong accumulator value is caught
    bad = trace.copy()
    bad[r, lay.mmcs_index_sum] = (int(bad[r, lay.mmcs_index_sum]) + 1) % F.p
    assert orc.check_constraints(inst, prep, bad, None) is not None
    # a wrong round value inside the permutation is caught
    bad = trace.copy()
    bad[3, lay.partial[5][1]] = (int(bad[3, lay.partial[5][1]]) + 1) % F.p
    assert orc.check_constraints(inst, prep, bad, None)[0] == 3
    # sponge continuation: capacity limbs of a chained row equal the previous row's output capacity
    chained = [i for i in range(1, n_ops) if prep[i, 4 * 2 + 2] == 1]
    assert chained
    i = chained[0]
    assert np.array_equal(trace[i, 8:16], trace[i - 1, lay.out_base + 8:lay.out_base + 16])


def test_bytecode_lowering_matches_direct_evaluation():
    """Random expression DAGs: the register-allocated program evaluated by the oracle interpreter (through
    orc_check_constraints) flags exactly the rows where a python big-int evaluation of the same DAG is non-zero."""
    orc = make_oracle("koala-bear")
    F = orc.field
    p = F.p
    rng = np.random.default_rng(12)
    width, height = 6, 16
    trace = F.rand(rng, (height, width))

    def eval_air(b):
        cols = [b.main(c) for c in range(width)] + [b.main(c, 1) for c in range(width)]
        e1 = cols[0] * cols[1] - cols[2]
        e2 = (cols[3] + cols[6]) * (cols[3] + cols[6]) * cols[4] - cols[5] * 7 + 3
        shared = cols[0] * cols[1]
        e3 = shared * cols[7] + shared - cols[8] * (1 - cols[9])
        for e in (e1, e2, e3):
            b.assert_zero(e)

    def direct(r):
        v = [int(x) for x in trace[r]] + [int(x) for x in trace[(r + 1) % height]]
        return [(v[0] * v[1] - v[2]) % p, ((v[3] + v[6]) ** 2 * v[4] - v[5] * 7 + 3) % p,
                (v[0] * v[1] * v[7] + v[0] * v[1] - v[8] * (1 - v[9])) % p]

    # make row 5 satisfy all three constraints by solving for columns 2, 5 (linear) and next-row column 8 - 6 = 2 of row 6
    inst = air_mod.build_instance("rand", eval_air, p, 4, width, 0, 0, air_mod.BusRegistry())
    res = orc.check_constraints(inst, None, trace, None)
    want = next(((r, k) for r in range(height) for k, val in enumerate(direct(r)) if val), None)
    assert res == want
    assert inst.constraints.n_constraints == 3 and inst.constraints.n_base_slots < 12


def test_specialized_kernels_are_current():
    """csrc/specialized_gen.cuh must have been generated from the current AIR programs (otherwise the library silently
    falls back to the interpreter)."""
    import os
    import re
    gen = importlib.import_module("scripts.gen_specialized") if False else None
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_specialized", os.path.join(root, "scripts", "gen_specialized.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    text = open(os.path.join(root, "plonky3-recursion_b200", "csrc", "specialized_gen.cuh")).read()
    have = set(re.findall(r"\{0x([0-9a-f]{16})ull, (\d), (\d+)u,", text))
    for fname in ("koala-bear", "baby-bear"):
        F = field_mod.get_field(fname)
        prm = p2mod.Poseidon2Params(F.field_id)
        buses = air_mod.BusRegistry()
        aw, apw = alu.widths(4, 3, 4)
        for inst in (air_mod.build_instance("alu", alu.make_eval(4, 3, 4, F.w), F.p, 8, aw, apw, 0, buses),
                     air_mod.build_instance("p2", p2air.make_eval(prm), F.p, 8, *p2air.widths(prm), 0, buses)):
            ins = g.monty_insns(F, inst.constraints)
            assert (f"{g.fnv1a(ins):016x}", str(F.field_id), str(ins.shape[0])) in have, (fname, inst.name)


def test_specialized_logup_kernels_are_current():
    """Same for the generated LogUp-trace kernels: hash over the lookup-input program and the lookup / interaction structure."""
    import importlib.util
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_specialized", os.path.join(root, "scripts", "gen_specialized.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    text = open(os.path.join(root, "plonky3-recursion_b200", "csrc", "specialized_gen.cuh")).read()
    have = set(re.findall(r"\{0x([0-9a-f]{16})ull, (\d), (\d+)u, k_logup_spec", text))
    F = field_mod.get_field("koala-bear")
    for coeff in (False, True):
        L = wl.synthetic_layer(F, 3, n_const=10, n_public=20, n_alu=60, n_perms=12, n_recompose=4, min_height=16,
                               recompose_coeff=coeff)
        for inst in L.insts:   # the layer's own instances must hit the registry, bus ids and all
            ins = g.monty_insns(F, inst.lookup_inputs)
            key = (f"{g.logup_hash(ins, inst.lookups, inst.interactions):016x}", str(F.field_id), str(ins.shape[0]))
            assert key in have, inst.name


@pytest.mark.parametrize("lanes", [1, 2])
def test_recompose_coeff_table_shape(lanes):
    """`recompose/coeff` (circuit-prover/src/air/recompose_air.rs:60-70,150-197): D main columns and 2 + 2D preprocessed
    columns per lane; 1 + D interactions per lane, one per lookup at this degree (no local constraints -> one quotient
    chunk -> budget 2), so 1 + lanes*(1 + D) permutation columns; coefficient tuples are (idx, v_i, 0, .., 0)."""
    F = field_mod.get_field("koala-bear")
    L = wl.synthetic_layer(F, 5, n_const=8, n_public=12, n_alu=80, n_perms=0, n_recompose=9, min_height=16,
                           recompose_coeff=True, recompose_lanes=lanes)
    inst, prep, trace = L.insts[-1], L.preps[-1], L.traces[-1]
    assert inst.name == "recompose/coeff"
    assert (inst.main_width, inst.prep_width) == (4 * lanes, 10 * lanes) == (trace.shape[1], prep.shape[1])
    assert inst.log_quotient_chunks == 0 and inst.aux_width == 1 + 5 * lanes and not inst.uses_next_row
    assert [n for _, _, n in inst.interactions] == [5] * (5 * lanes)
    # hint-output operations (even) carry read counts, the others multiplicity 0; padding rows are all zero
    flat = prep.reshape(-1, 10)
    assert (flat[1:9:2, 3::2] == 0).all() and (flat[9:] == 0).all()
    assert flat[0:9:2, 3::2].sum() > 0


@pytest.mark.parametrize("field", ["koala-bear", "baby-bear"])
def test_alu_schedule_slots_reproduce_the_table(field):
    """AluTableOps (what p3r_alu_ops carries: schedule slots + operand values) determines the ALU main trace: a numpy
    restatement of k_alu_table_fill — rows independent, lane-0 accumulator = out of the previous row's lane-0 slot — rebuilds
    the matrix of AluAir::trace_to_matrix (alu.build_tables) exactly."""
    F = field_mod.get_field(field)
    L = wl.synthetic_layer(F, 9, n_const=10, n_public=30, n_alu=250, n_perms=10, n_recompose=4, min_height=16)
    (idx, t), = L.alu_ops.items()
    want = L.traces[idx]
    d, lanes, k_max = t.d, t.lanes, t.k_max
    assert (d, k_max) == (4, 4) and (t.slot_kind >= 2).any()
    assert not ((t.slot_kind >= 2) & (np.arange(t.slot_kind.size) % lanes != 0)).any()   # packed runs sit in lane 0
    H = want.shape[0]
    got = np.zeros_like(want)
    extra, num_int = lanes * 4 * d, (k_max - 1) // 2
    ac_base = extra + num_int * d
    bsq = ac_base + 2 * (k_max - 1) * d
    mul = lambda x, y: F.ext_mul([int(v) for v in x], [int(v) for v in y])
    add = lambda x, y: [(int(u) + int(v)) % F.p for u, v in zip(x, y)]
    sub = lambda x, y: [(int(u) - int(v)) % F.p for u, v in zip(x, y)]
    for row in range(H):
        for lane in range(lanes):
            s = row * lanes + lane
            if s >= t.slot_kind.size or t.slot_kind[s] == 0:
                continue
            k, first = int(t.slot_kind[s]), int(t.slot_first[s])
            m = lane * 4 * d
            got[row, m:m + 3 * d] = t.values[first, :3].reshape(-1)
            got[row, m + 3 * d:m + 4 * d] = t.values[first + k - 1, 3]
            if k >= 2 and lane == 0:
                acc = [0] * d
                if row > 0 and t.slot_kind[(row - 1) * lanes]:
                    ps = (row - 1) * lanes
                    acc = list(t.values[int(t.slot_first[ps]) + int(t.slot_kind[ps]) - 1, 3])
                b = t.values[first, 1]
                step = 0
                for si in range(num_int):
                    i0 = first + step
                    acc = sub(add(mul(acc, b), t.values[i0, 2]), t.values[i0, 0])
                    step += 1
                    if i0 + 1 < first + k:
                        acc = sub(add(mul(acc, b), t.values[i0 + 1, 2]), t.values[i0 + 1, 0])
                        step += 1
                    got[row, extra + si * d:extra + (si + 1) * d] = acc
                for tt in range(1, k):
                    got[row, ac_base + 2 * (tt - 1) * d:ac_base + (2 * (tt - 1) + 1) * d] = t.values[first + tt, 0]
                    got[row, ac_base + (2 * (tt - 1) + 1) * d:ac_base + 2 * tt * d] = t.values[first + tt, 2]
                got[row, bsq:bsq + d] = mul(b, b)
    assert np.array_equal(got, want)
