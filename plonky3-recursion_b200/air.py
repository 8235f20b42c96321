"""AirInstance: one table as the C ABI sees it (p3r_instance_desc), built from an evaluated AirBuilder.

Mirrors what `ProverData::from_airs_and_degrees` derives per AIR in the reference
(/root/reference circuit-prover/src/batch_stark_prover.rs:915-949): symbolic constraints, `Lookups::from_air`,
`get_log_num_quotient_chunks`, `pack_same_bus(gadget, budget)` with budget = 2^log_chunks + 1.
"""
from __future__ import annotations

from .symbolic import (AirBuilder, compile_constraints, compile_outputs, log_quotient_chunks, logup_constraints,
                       pack_same_bus)


class BusRegistry:
    """Stable global bus ids in iteration order (recursion/src/verifier/batch_stark.rs:1055-1083)."""

    def __init__(self):
        self.ids = {}

    def get(self, name: str) -> int:
        return self.ids.setdefault(name, len(self.ids))


class AirInstance:
    def __init__(self, name, log_height, main_width, prep_width, n_public, log_qc, uses_next_row, constraints,
                 lookup_inputs, lookups, interactions, air_only_constraints=None):
        self.name = name
        self.log_height, self.main_width, self.prep_width, self.n_public = log_height, main_width, prep_width, n_public
        self.log_quotient_chunks, self.uses_next_row = log_qc, uses_next_row
        self.constraints, self.lookup_inputs = constraints, lookup_inputs
        self.lookups, self.interactions = lookups, interactions      # [(bus, first, n)], [(mult_out, elem_first, n_elems)]
        self.air_only_constraints = air_only_constraints            # Program without LogUp (for check_constraints)

    @property
    def aux_width(self):
        return len(self.lookups) + 1 if self.lookups else 0


def _clone_builder(make_builder):
    b = make_builder()
    return b


def build_instance(name: str, eval_air, p: int, log_height: int, main_width: int, prep_width: int, n_public: int,
                   buses: BusRegistry) -> AirInstance:
    """eval_air(builder) evaluates the AIR (constraints + push_interaction) into a fresh AirBuilder."""

    def fresh():
        b = AirBuilder(p, main_width, prep_width, n_public)
        eval_air(b)
        return b

    # pass 1: AIR + unpacked lookups -> log_chunks -> budget (batch_stark_prover.rs:925-941)
    b1 = fresh()
    n_air_base = len(b1.base_constraints)
    air_only = compile_constraints(b1) if (b1.base_constraints or b1.ext_constraints) else None
    unpacked = [[it] for it in b1.interactions]
    logup_constraints(b1, unpacked)
    log_chunks = log_quotient_chunks(b1)
    budget = (1 << log_chunks) + 1

    # pass 2: packed lookups
    b = fresh()
    groups = pack_same_bus(b.interactions, budget)
    logup_constraints(b, groups)
    log_qc = log_quotient_chunks(b)
    assert log_qc == log_chunks, (log_qc, log_chunks)
    assert len(b.base_constraints) == n_air_base
    constraints = compile_constraints(b)

    outs, lookups, interactions = [], [], []
    for g in groups:
        lookups.append((buses.get(g[0].bus), len(interactions), len(g)))
        for it in g:
            mult_out = len(outs)
            outs.append(it.mult)
            elem_first = len(outs)
            outs.extend(it.fields)
            interactions.append((mult_out, elem_first, len(it.fields)))
    lookup_inputs = compile_outputs(outs) if outs else None
    return AirInstance(name, log_height, main_width, prep_width, n_public, log_qc, b.uses_next_row, constraints,
                       lookup_inputs, lookups, interactions, air_only)
