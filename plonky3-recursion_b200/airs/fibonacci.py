"""Two-column Fibonacci test AIR with three public values (a0, b0, last b): exercises is_first_row / is_transition /
is_last_row selectors, public values, and the no-lookup / no-preprocessed proof shape (the analogue of the reference's
test fixtures such as MulAir, /root/reference recursion/tests/common/mod.rs:45-160)."""
from __future__ import annotations

import numpy as np


def eval_air(b):
    a, bb = b.main(0), b.main(1)
    an, bn = b.main(0, 1), b.main(1, 1)
    b.when_first_row().assert_eq(a, b.public(0))
    b.when_first_row().assert_eq(bb, b.public(1))
    tr = b.when_transition()
    tr.assert_eq(an, bb)
    tr.assert_eq(bn, a + bb)
    b.when_last_row().assert_eq(bb, b.public(2))


def trace(p: int, log_n: int, a0: int = 0, b0: int = 1):
    n = 1 << log_n
    t = np.zeros((n, 2), dtype=np.uint32)
    a, b = a0 % p, b0 % p
    for i in range(n):
        t[i] = (a, b)
        a, b = b, (a + b) % p
    return t, np.array([a0 % p, b0 % p, int(t[-1, 1])], dtype=np.uint32)
