"""A wide single-table AIR shaped like the base layer of `recursive_keccak` (recursion/examples/recursive_keccak.rs:22-24,513-530:
`KeccakAir`, ~2 600 columns, rows = 24 per permutation padded to a power of two, proved by `p3_uni_stark::prove`).

KeccakAir itself lives in p3-keccak-air (crates.io, not in the reference tree), so the constraints here are SYNTHETIC with the
same cost profile: `width` columns, about one low-degree constraint per column, degree 3 overall (two quotient chunks), next-row
transitions, three public values. Per row:
    c_0' = c_0 + 1            (row counter, starts at public[0])
    c_1' = c_1 * public[1]    (geometric sequence, starts at 1)
    c_2  = c_0 * c_1
    c_j  = c_{j-1} * c_{j-2} + c_{j-3}                         for j = 3 .. width-1, j % 8 != 0
    c_j  = c_{j-1} * c_{j-2} * c_{j-3} + c_{j-4}   (degree 3)   for j % 8 == 0, j >= 8
    last row: c_{width-1} = public[2]
"""
from __future__ import annotations

import numpy as np

N_PUBLIC = 3


def make_eval(width: int):
    assert width >= 9

    def eval_air(b):
        c = [b.main(j) for j in range(width)]
        first, trans, last = b.when_first_row(), b.when_transition(), b.when_last_row()
        first.assert_eq(c[0], b.public(0))
        first.assert_eq(c[1], 1)
        trans.assert_eq(b.main(0, 1), c[0] + 1)
        trans.assert_eq(b.main(1, 1), c[1] * b.public(1))
        b.assert_eq(c[2], c[0] * c[1])
        for j in range(3, width):
            if j % 8 == 0 and j >= 8:
                b.assert_eq(c[j], c[j - 1] * c[j - 2] * c[j - 3] + c[j - 4])
            else:
                b.assert_eq(c[j], c[j - 1] * c[j - 2] + c[j - 3])
        last.assert_eq(c[width - 1], b.public(2))
    return eval_air


def trace(p: int, width: int, log_n: int, start: int = 5, ratio: int = 3):
    """(n, width) canonical trace and the public values [start, ratio, last row's last column]."""
    n = 1 << log_n
    P = np.uint64(p)
    t = np.zeros((n, width), dtype=np.uint64)
    t[:, 0] = (np.arange(n, dtype=np.uint64) + np.uint64(start)) % P
    g = np.ones(n, dtype=np.uint64)
    for i in range(1, n):
        g[i] = g[i - 1] * np.uint64(ratio) % P
    t[:, 1] = g
    t[:, 2] = t[:, 0] * t[:, 1] % P
    for j in range(3, width):
        if j % 8 == 0 and j >= 8:
            t[:, j] = (t[:, j - 1] * t[:, j - 2] % P * t[:, j - 3] + t[:, j - 4]) % P
        else:
            t[:, j] = (t[:, j - 1] * t[:, j - 2] + t[:, j - 3]) % P
    return t.astype(np.uint32), np.array([start % p, ratio % p, int(t[-1, width - 1])], dtype=np.uint32)
