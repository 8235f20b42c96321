"""Symbolic AIR builder -> p3r constraint bytecode (include/p3r.h).

Host-side mirror of what the Rust integration does once per circuit shape: take the (base, ext)
SymbolicExpression DAGs of an AIR (p3_batch_stark::symbolic::get_symbolic_constraints, the call made at
/root/reference recursion/src/traits/air.rs:160), append the LogUp constraints, and lower them to the
register bytecode the GPU quotient kernel interprets. Node kinds follow
circuit/src/symbolic/compiler.rs:86-118,183-189; degrees follow p3-air's `degree_multiple`
(variables 1, is_first_row/is_last_row 1, is_transition/public/challenge/constant 0).

Expressions are hash-consed per `ExprCtx` so shared sub-expressions become one instruction (the CSE cache
of circuit/src/symbolic/compiler.rs:47,143). All constants are canonical integers here; Montgomery
encoding happens when the program is serialised for the C ABI.
"""
from __future__ import annotations

import numpy as np

# opcodes (include/p3r.h p3r_opcode)
OP_B_MAIN, OP_B_PREP, OP_B_PUB, OP_B_SEL, OP_B_CONST, OP_B_ADD, OP_B_SUB, OP_B_MUL, OP_B_NEG = range(9)
OP_E_PERM, OP_E_CHAL, OP_E_PVAL, OP_E_CONST, OP_E_FROMB = 16, 17, 18, 19, 20
OP_E_ADD, OP_E_SUB, OP_E_MUL, OP_E_NEG, OP_E_MULB, OP_E_ADDB, OP_E_SUBB = 21, 22, 23, 24, 25, 26, 27
OP_ASSERT_B, OP_ASSERT_E, OP_OUT_B = 32, 33, 34

SEL_FIRST, SEL_LAST, SEL_TRANSITION = 0, 1, 2

_LEAF_B = {"main": OP_B_MAIN, "prep": OP_B_PREP, "pub": OP_B_PUB, "sel": OP_B_SEL, "const": OP_B_CONST}
_BIN_B = {"add": OP_B_ADD, "sub": OP_B_SUB, "mul": OP_B_MUL}
_LEAF_E = {"perm": OP_E_PERM, "chal": OP_E_CHAL, "pval": OP_E_PVAL, "econst": OP_E_CONST}
_BIN_E = {"eadd": OP_E_ADD, "esub": OP_E_SUB, "emul": OP_E_MUL}
_MIX_E = {"emulb": OP_E_MULB, "eaddb": OP_E_ADDB, "esubb": OP_E_SUBB}


class Expr:
    __slots__ = ("ctx", "kind", "args", "deg", "ext", "uid")

    def __init__(self, ctx, kind, args, deg, ext, uid):
        self.ctx, self.kind, self.args, self.deg, self.ext, self.uid = ctx, kind, args, deg, ext, uid

    # -- arithmetic; mixes base/ext transparently -------------------------------------------------
    def _coerce(self, o):
        return o if isinstance(o, Expr) else self.ctx.const(o)

    def __add__(self, o):
        return self.ctx.add(self, self._coerce(o))

    __radd__ = __add__

    def __sub__(self, o):
        return self.ctx.sub(self, self._coerce(o))

    def __rsub__(self, o):
        return self.ctx.sub(self._coerce(o), self)

    def __mul__(self, o):
        return self.ctx.mul(self, self._coerce(o))

    __rmul__ = __mul__

    def __neg__(self):
        return self.ctx.neg(self)

    def is_const(self):
        return self.kind == "const"

    def __repr__(self):
        return f"<{self.kind}{self.args if self.kind in _LEAF_B or self.kind in _LEAF_E else ''} d{self.deg}>"


class ExprCtx:
    """Hash-consing factory for one AIR over the field with modulus p."""

    def __init__(self, p: int):
        self.p = p
        self._table = {}
        self._n = 0

    def _mk(self, kind, args, deg, ext):
        key = (kind, tuple(a.uid if isinstance(a, Expr) else a for a in args))
        e = self._table.get(key)
        if e is None:
            e = Expr(self, kind, tuple(args), deg, ext, self._n)
            self._n += 1
            self._table[key] = e
        return e

    # leaves
    def main(self, col, off=0):
        return self._mk("main", (int(col), int(off)), 1, False)

    def prep(self, col, off=0):
        return self._mk("prep", (int(col), int(off)), 1, False)

    def pub(self, i):
        return self._mk("pub", (int(i),), 0, False)

    def sel(self, which):
        return self._mk("sel", (int(which),), 0 if which == SEL_TRANSITION else 1, False)

    def const(self, v):
        return self._mk("const", (int(v) % self.p,), 0, False)

    def perm(self, col, off=0):
        return self._mk("perm", (int(col), int(off)), 1, True)

    def chal(self, i):
        return self._mk("chal", (int(i),), 0, True)

    def pval(self, i):
        return self._mk("pval", (int(i),), 0, True)

    def econst(self, coeffs):
        return self._mk("econst", tuple(int(c) % self.p for c in coeffs), 0, True)

    def lift(self, b):
        return b if b.ext else self._mk("fromb", (b,), b.deg, True)

    # ops with light constant folding (keeps the bytecode small; does not change the constraint values)
    def add(self, a, b):
        if a.ext or b.ext:
            if not a.ext:
                a, b = b, a
            if not b.ext:
                if b.is_const() and b.args[0] == 0:
                    return a
                return self._mk("eaddb", (a, b), max(a.deg, b.deg), True)
            return self._mk("eadd", (a, b), max(a.deg, b.deg), True)
        if a.is_const() and b.is_const():
            return self.const(a.args[0] + b.args[0])
        if a.is_const() and a.args[0] == 0:
            return b
        if b.is_const() and b.args[0] == 0:
            return a
        return self._mk("add", (a, b), max(a.deg, b.deg), False)

    def sub(self, a, b):
        if a.ext or b.ext:
            if a.ext and not b.ext:
                if b.is_const() and b.args[0] == 0:
                    return a
                return self._mk("esubb", (a, b), max(a.deg, b.deg), True)
            if not a.ext:
                a = self.lift(a)
            return self._mk("esub", (a, b), max(a.deg, b.deg), True)
        if a.is_const() and b.is_const():
            return self.const(a.args[0] - b.args[0])
        if b.is_const() and b.args[0] == 0:
            return a
        return self._mk("sub", (a, b), max(a.deg, b.deg), False)

    def mul(self, a, b):
        if a.ext or b.ext:
            if not a.ext:
                a, b = b, a
            if not b.ext:
                if b.is_const() and b.args[0] == 1:
                    return a
                return self._mk("emulb", (a, b), a.deg + b.deg, True)
            return self._mk("emul", (a, b), a.deg + b.deg, True)
        if a.is_const() and b.is_const():
            return self.const(a.args[0] * b.args[0])
        for x, y in ((a, b), (b, a)):
            if x.is_const():
                if x.args[0] == 0:
                    return x
                if x.args[0] == 1:
                    return y
        return self._mk("mul", (a, b), a.deg + b.deg, False)

    def neg(self, a):
        if a.ext:
            return self._mk("eneg", (a,), a.deg, True)
        if a.is_const():
            return self.const(-a.args[0])
        return self._mk("neg", (a,), a.deg, False)


class Interaction:
    """One `push_interaction(bus, fields, Count::bounded(mult, _))` (p3-lookup InteractionBuilder)."""

    def __init__(self, bus: str, fields, mult):
        self.bus, self.fields, self.mult = bus, list(fields), mult


class AirBuilder:
    """Collects constraints + interactions; mirrors p3-air's AirBuilder filtering API."""

    def __init__(self, p: int, main_width: int, prep_width: int = 0, n_public: int = 0):
        self.ctx = ExprCtx(p)
        self.p = p
        self.main_width, self.prep_width, self.n_public = main_width, prep_width, n_public
        self.base_constraints = []
        self.ext_constraints = []
        self.interactions = []
        self._cond = None
        self.uses_next_row = False

    # variables
    def main(self, col, off=0):
        assert 0 <= col < self.main_width, (col, self.main_width)
        if off:
            self.uses_next_row = True
        return self.ctx.main(col, off)

    def prep(self, col, off=0):
        assert 0 <= col < self.prep_width, (col, self.prep_width)
        return self.ctx.prep(col, off)

    def local(self):
        return [self.main(c, 0) for c in range(self.main_width)]

    def next(self):
        return [self.main(c, 1) for c in range(self.main_width)]

    def prep_local(self):
        return [self.prep(c, 0) for c in range(self.prep_width)]

    def prep_next(self):
        return [self.prep(c, 1) for c in range(self.prep_width)]

    def public(self, i):
        return self.ctx.pub(i)

    def const(self, v):
        return self.ctx.const(v)

    def is_first_row(self):
        return self.ctx.sel(SEL_FIRST)

    def is_last_row(self):
        return self.ctx.sel(SEL_LAST)

    def is_transition(self):
        return self.ctx.sel(SEL_TRANSITION)

    # filtering
    def when(self, cond):
        b = _Filtered(self, cond if self._cond is None else self._cond * cond)
        return b

    def when_first_row(self):
        return self.when(self.is_first_row())

    def when_last_row(self):
        return self.when(self.is_last_row())

    def when_transition(self):
        return self.when(self.is_transition())

    def assert_zero(self, e, cond=None):
        e = e if isinstance(e, Expr) else self.ctx.const(e)
        if cond is not None:
            e = cond * e
        (self.ext_constraints if e.ext else self.base_constraints).append(e)

    def assert_eq(self, a, b, cond=None):
        a = a if isinstance(a, Expr) else self.ctx.const(a)
        self.assert_zero(a - b, cond)

    def assert_bool(self, x, cond=None):
        self.assert_zero(x * (x - 1), cond)

    def push_interaction(self, bus, fields, mult):
        fields = [f if isinstance(f, Expr) else self.ctx.const(f) for f in fields]
        mult = mult if isinstance(mult, Expr) else self.ctx.const(mult)
        self.interactions.append(Interaction(bus, fields, mult))


class _Filtered:
    def __init__(self, b, cond):
        self.b, self.cond = b, cond

    def when(self, cond):
        return _Filtered(self.b, self.cond * cond)

    def when_transition(self):
        return self.when(self.b.is_transition())

    def assert_zero(self, e):
        self.b.assert_zero(e, self.cond)

    def assert_eq(self, a, b):
        self.b.assert_eq(a, b, self.cond)

    def assert_bool(self, x):
        self.b.assert_bool(x, self.cond)


# ---------------------------------------------------------------------------------------------------
# LogUp: pack interactions into lookups (fraction columns) and emit the LogUp constraints.
# Layout restated in-tree at recursion/src/verifier/batch_stark.rs:902-912 (col 0 accumulator, col c+1
# fraction column of lookup c, one terminal per AIR). The exact constraint list of p3-lookup 0.6 is
# [P3-EXT]; the list below is this framework's definition (DESIGN.md "LogUp").
# ---------------------------------------------------------------------------------------------------
def pack_same_bus(interactions, budget):
    """Greedy `pack_same_bus(gadget, budget)` (circuit-prover/src/batch_stark_prover.rs:925-941): put
    consecutive interactions of one bus into one fraction column while the column's constraint degree
    frac*prod(den) - sum(mult*prod(other den)) stays <= budget."""
    groups = []
    for it in interactions:
        dden = max([f.deg for f in it.fields] + [0])

        def degree(group):
            dens = [max([f.deg for f in g.fields] + [0]) for g in group]
            lhs = 1 + sum(dens)
            rhs = max(g.mult.deg + sum(dens) - d for g, d in zip(group, dens))
            return max(lhs, rhs)

        if groups and groups[-1][0].bus == it.bus and degree(groups[-1] + [it]) <= budget:
            groups[-1].append(it)
        else:
            if degree([it]) > budget:
                raise ValueError(f"interaction degree {degree([it])} exceeds budget {budget} (dden={dden})")
            groups.append([it])
    return groups


# [P3-EXT] LogUp denominator convention (include/p3r.h p3r_conventions): prefix + s * sum_k beta^{e(k)} f_k with s = -1 if
# `negate`, e(k) = first_power + (n - 1 - k if descending else k). Must match Context.set_conventions / Oracle.set_conventions.
LOGUP_CONVENTIONS = dict(logup_negate=0, logup_first_power=0, logup_descending=0)


def logup_constraints(b: AirBuilder, groups):
    """Append the LogUp constraints for `groups` (list of lists of Interaction) to builder `b`."""
    ctx = b.ctx
    if not groups:
        return
    cv = LOGUP_CONVENTIONS
    fracs = []
    for c, group in enumerate(groups):
        prefix, beta = ctx.chal(2 * c), ctx.chal(2 * c + 1)
        dens = []
        for it in group:
            n = len(it.fields)
            pows = [None]                    # pows[e] = beta^e as an expression (None = 1)
            for _ in range(n + 1):
                pows.append(beta if pows[-1] is None else pows[-1] * beta)
            den = prefix
            for k, f in enumerate(it.fields):
                e = cv["logup_first_power"] + ((n - 1 - k) if cv["logup_descending"] else k)
                term = f if pows[e] is None else pows[e] * f    # default convention: the same nodes as ever (program hash)
                den = (den - term) if cv["logup_negate"] else (den + term)
            dens.append(den)
        frac = ctx.perm(c + 1, 0)
        fracs.append(frac)
        lhs = frac
        for d in dens:
            lhs = lhs * d
        rhs = None
        for j, it in enumerate(group):
            t = ctx.lift(it.mult)
            for k, d in enumerate(dens):
                if k != j:
                    t = t * d
            rhs = t if rhs is None else rhs + t
        b.assert_zero(lhs - rhs)
    acc, acc_next, terminal = ctx.perm(0, 0), ctx.perm(0, 1), ctx.pval(0)
    total = fracs[0]
    for f in fracs[1:]:
        total = total + f
    b.assert_zero(ctx.lift(b.is_first_row()) * acc)
    b.assert_zero(ctx.lift(b.is_transition()) * (acc_next - acc - total))
    b.assert_zero(ctx.lift(b.is_last_row()) * (terminal - acc - total))


# ---------------------------------------------------------------------------------------------------
# Lowering: DAG -> linear bytecode with slot (register) allocation by last use.
# ---------------------------------------------------------------------------------------------------
class Program:
    def __init__(self, insns, n_base_slots, n_ext_slots, ext_consts, n_constraints=0, n_outputs=0):
        self.insns = np.asarray(insns, dtype=np.uint32).reshape(-1, 4)  # canonical immediates
        self.n_base_slots, self.n_ext_slots = n_base_slots, n_ext_slots
        self.ext_consts = np.asarray(ext_consts, dtype=np.uint32).reshape(-1, 4)
        self.n_constraints, self.n_outputs = n_constraints, n_outputs


def _lower(roots, sinks):
    """roots: list of Expr; sinks: list of (opcode, dst_index) parallel to roots."""
    order, seen = [], set()
    for r in roots:  # iterative post-order
        stack = [(r, False)]
        while stack:
            e, done = stack.pop()
            if done:
                order.append(e)
                continue
            if e.uid in seen:
                continue
            seen.add(e.uid)
            stack.append((e, True))
            for a in e.args:
                if isinstance(a, Expr) and a.uid not in seen:
                    stack.append((a, False))
    # schedule: nodes in post-order, each root's sink right after the root is available
    pos = {e.uid: i for i, e in enumerate(order)}
    sink_at = {}
    for r, s in zip(roots, sinks):
        sink_at.setdefault(pos[r.uid], []).append((r, s))
    last_use = {}
    for i, e in enumerate(order):
        for a in e.args:
            if isinstance(a, Expr):
                last_use[a.uid] = i
    for r in roots:
        last_use[r.uid] = max(last_use.get(r.uid, -1), pos[r.uid])
    free = {False: [], True: []}
    nslots = {False: 0, True: 0}
    slot = {}
    econsts, econst_idx = [], {}
    insns = []

    def alloc(ext):
        if free[ext]:
            return free[ext].pop()
        nslots[ext] += 1
        return nslots[ext] - 1

    released = set()

    def release(x):
        if x.uid not in released:
            released.add(x.uid)
            free[x.ext].append(slot[x.uid])

    for i, e in enumerate(order):
        k = e.kind
        argslots = [slot[a.uid] for a in e.args if isinstance(a, Expr)]
        # operands dying here are released BEFORE dst is allocated: dst may alias a source slot, which both
        # interpreters allow (operands are read into temporaries before the destination is written).
        for a in e.args:
            if isinstance(a, Expr) and last_use[a.uid] == i:
                release(a)
        d = alloc(e.ext)
        slot[e.uid] = d
        if k in ("main", "prep", "perm"):
            op = _LEAF_B.get(k, _LEAF_E.get(k))
            insns.append((op, d, e.args[0], e.args[1]))
        elif k in ("pub", "sel", "const", "chal", "pval"):
            op = _LEAF_B.get(k, _LEAF_E.get(k))
            insns.append((op, d, e.args[0], 0))
        elif k == "econst":
            if e.args not in econst_idx:
                econst_idx[e.args] = len(econsts)
                econsts.append(e.args)
            insns.append((OP_E_CONST, d, econst_idx[e.args], 0))
        elif k == "fromb":
            insns.append((OP_E_FROMB, d, argslots[0], 0))
        elif k in _BIN_B:
            insns.append((_BIN_B[k], d, argslots[0], argslots[1]))
        elif k == "neg":
            insns.append((OP_B_NEG, d, argslots[0], 0))
        elif k in _BIN_E:
            insns.append((_BIN_E[k], d, argslots[0], argslots[1]))
        elif k in _MIX_E:
            insns.append((_MIX_E[k], d, argslots[0], argslots[1]))
        elif k == "eneg":
            insns.append((OP_E_NEG, d, argslots[0], 0))
        else:
            raise ValueError(k)
        for r, (op, idx) in sink_at.get(i, []):
            insns.append((op, idx, slot[r.uid], 0))
        if last_use.get(e.uid, i) <= i:
            release(e)  # a root that nothing reads afterwards
    return insns, max(nslots[False], 1), max(nslots[True], 1), econsts


def compile_constraints(b: AirBuilder) -> Program:
    """Base constraints first, then extension constraints (recursion/src/traits/air.rs:170-181)."""
    roots = list(b.base_constraints) + list(b.ext_constraints)
    sinks = [(OP_ASSERT_B, i) for i in range(len(b.base_constraints))] + [
        (OP_ASSERT_E, len(b.base_constraints) + i) for i in range(len(b.ext_constraints))
    ]
    insns, nb, ne, ec = _lower(roots, sinks)
    return Program(insns, nb, ne, ec, n_constraints=len(roots))


def compile_outputs(ctx_exprs) -> Program:
    roots = list(ctx_exprs)
    for r in roots:
        assert not r.ext, "lookup inputs are base-field expressions"
    sinks = [(OP_OUT_B, i) for i in range(len(roots))]
    insns, nb, ne, ec = _lower(roots, sinks)
    return Program(insns, nb, ne, ec, n_outputs=len(roots))


def max_constraint_degree(b: AirBuilder) -> int:
    return max([e.deg for e in b.base_constraints + b.ext_constraints] + [0])


def log_quotient_chunks(b: AirBuilder) -> int:
    """p3 get_log_quotient_degree: log2_ceil(max(constraint_degree, 2) - 1) (non-ZK)."""
    d = max(max_constraint_degree(b), 2)
    return int(np.ceil(np.log2(d - 1))) if d > 2 else 0
