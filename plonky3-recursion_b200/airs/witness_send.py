"""WitnessSendAir: the AIR behind ConstAir and PublicAir.

Reference: /root/reference circuit-prover/src/air/public_air.rs:43-230 (WitnessSendAir), const_air.rs:52-158 (ConstAir is the
single-lane case). Main trace: `lanes * D` value coordinates; preprocessed: per lane `[multiplicity, witness_idx]`
(column_layout.rs WITNESS_LOOKUP_PREP_COL_MAP). No local constraints; one `WitnessChecks` send per lane
`(witness_idx, value[0..D])` with the preprocessed multiplicity (public_air.rs:202-230).
Shape goldens (air/shape_golden.rs:33-45): Const D4 -> (4, 2); Public D4 x 2 lanes -> (8, 4).
"""
from __future__ import annotations

import numpy as np

PREP_LANE_WIDTH = 2  # [multiplicity, witness_idx]


def widths(d: int, lanes: int = 1):
    return lanes * d, lanes * PREP_LANE_WIDTH


def prep_lane_width(d: int, coeff_lookups: bool = False) -> int:
    """[.., ..] plus, on the `recompose/coeff` table, `(coeff_i_idx, coeff_i_mult)` per coefficient
    (recompose_air.rs:60-70, `preprocessed_lane_width_for`)."""
    return PREP_LANE_WIDTH + (2 * d if coeff_lookups else 0)


def make_eval(d: int, lanes: int = 1, idx_first: bool = False, coeff_lookups: bool = False):
    """idx_first=False: [multiplicity, witness_idx] (Const/Public, column_layout.rs); idx_first=True: [output_idx, out_mult]
    (RecomposeAir, circuit-prover/src/air/recompose_air.rs:150-175). coeff_lookups=True is the `recompose/coeff` table
    (recompose_air.rs:175-197): per lane, after the output interaction, D more interactions
    `[coeff_i_idx, v_i, 0, .., 0]` with multiplicity `coeff_i_mult`, in coefficient order."""
    assert idx_first or not coeff_lookups
    plw = prep_lane_width(d, coeff_lookups)

    def eval_air(b):
        for lane in range(lanes):
            mult = b.prep(lane * plw + (1 if idx_first else 0))
            idx = b.prep(lane * plw + (0 if idx_first else 1))
            fields = [idx] + [b.main(lane * d + j) for j in range(d)]
            b.push_interaction("WitnessChecks", fields, mult)
            if coeff_lookups:
                for i in range(d):
                    cidx = b.prep(lane * plw + PREP_LANE_WIDTH + 2 * i)
                    cmult = b.prep(lane * plw + PREP_LANE_WIDTH + 2 * i + 1)
                    b.push_interaction("WitnessChecks", [cidx, b.main(lane * d + i)] + [b.const(0)] * (d - 1), cmult)

    return eval_air


def trace_to_matrix(values: np.ndarray, d: int, lanes: int, min_height: int) -> np.ndarray:
    """values: (num_ops, d) canonical. Lane packing + zero padding to max(pow2, min_height)
    (public_air.rs:128-170, const_air.rs:91-126)."""
    values = np.asarray(values, dtype=np.uint32).reshape(-1, d)
    num_ops = values.shape[0]
    rows = -(-num_ops // lanes) if num_ops else 0
    height = max(min_height, 1 << max(rows - 1, 0).bit_length())
    out = np.zeros((height, lanes * d), dtype=np.uint32)
    flat = out.reshape(height * lanes, d)
    flat[:num_ops] = values
    return out


def preprocessed_matrix(mults: np.ndarray, idxs: np.ndarray, lanes: int, min_height: int, idx_first: bool = False,
                        coeff_idxs: np.ndarray | None = None, coeff_mults: np.ndarray | None = None) -> np.ndarray:
    """Per op (multiplicity, D-scaled witness index), canonical; padding rows have multiplicity 0
    (circuit-prover/src/common.rs:226-287). coeff_idxs / coeff_mults, both (num_ops, D): the `recompose/coeff` columns
    `(coeff_i_idx, coeff_i_mult)` appended per lane (circuit-prover/src/batch_stark_prover/recompose.rs:308-352)."""
    mults = np.asarray(mults, dtype=np.uint32)
    idxs = np.asarray(idxs, dtype=np.uint32)
    num_ops = mults.shape[0]
    rows = -(-num_ops // lanes) if num_ops else 0
    height = max(min_height, 1 << max(rows - 1, 0).bit_length())
    plw = PREP_LANE_WIDTH
    if coeff_idxs is not None:
        coeff_idxs = np.asarray(coeff_idxs, dtype=np.uint32).reshape(num_ops, -1)
        coeff_mults = np.asarray(coeff_mults, dtype=np.uint32).reshape(num_ops, -1)
        plw += 2 * coeff_idxs.shape[1]
    out = np.zeros((height, lanes * plw), dtype=np.uint32)
    flat = out.reshape(height * lanes, plw)
    flat[:num_ops, 1 if idx_first else 0] = mults
    flat[:num_ops, 0 if idx_first else 1] = idxs
    if coeff_idxs is not None:
        flat[:num_ops, PREP_LANE_WIDTH::2] = coeff_idxs
        flat[:num_ops, PREP_LANE_WIDTH + 1::2] = coeff_mults
    return out
