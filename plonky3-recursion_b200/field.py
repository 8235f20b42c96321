"""Field parameters and small numpy helpers (host side; canonical residues unless named *_monty).

KoalaBear p = 0x7f000001, BabyBear p = 0x78000001 (/root/reference circuit-prover/src/batch_stark_prover/tests.rs:704-705,
731-732); W is the binomial constant x^4 = W (circuit-prover/src/field_params.rs:34-41), GENERATOR the coset shift
(recursion/src/pcs/fri/verifier.rs:960). W/GENERATOR values are the p3 ones as recalled ([P3-EXT]) and
checked here only for the properties that matter (generator order p-1, W a quartic non-residue).
"""
from __future__ import annotations

import numpy as np

KOALABEAR = 0
BABYBEAR = 1


class Field:
    def __init__(self, field_id, name, p, generator, w, two_adicity):
        self.field_id, self.name, self.p, self.generator, self.w, self.two_adicity = field_id, name, p, generator, w, two_adicity
        self.r_inv = pow(1 << 32, p - 2, p)

    # Montgomery form used at the C ABI: monty(x) = x * 2^32 mod p
    def to_monty(self, a):
        a = np.asarray(a, dtype=np.uint64)
        return ((a << np.uint64(32)) % np.uint64(self.p)).astype(np.uint32)

    def from_monty(self, a):
        a = np.asarray(a, dtype=np.uint64)
        return ((a * np.uint64(self.r_inv)) % np.uint64(self.p)).astype(np.uint32)

    def inv(self, a: int) -> int:
        return pow(int(a) % self.p, self.p - 2, self.p)

    def two_adic_generator(self, bits: int) -> int:
        assert bits <= self.two_adicity
        return pow(self.generator, (self.p - 1) >> bits, self.p)

    # binomial extension helpers on python ints (tiny; used by tests / public-value plumbing)
    def ext_mul(self, a, b):
        t = [0] * 7
        for i in range(4):
            for j in range(4):
                t[i + j] = (t[i + j] + a[i] * b[j]) % self.p
        return [(t[i] + (self.w * t[i + 4] if i < 3 else 0)) % self.p for i in range(4)]

    def rand(self, rng: np.random.Generator, shape):
        return rng.integers(0, self.p, size=shape, dtype=np.uint64).astype(np.uint32)


FIELDS = {
    KOALABEAR: Field(KOALABEAR, "koala-bear", 0x7F000001, 3, 3, 24),
    BABYBEAR: Field(BABYBEAR, "baby-bear", 0x78000001, 31, 11, 27),
}


def get_field(name_or_id) -> Field:
    if isinstance(name_or_id, Field):
        return name_or_id
    if isinstance(name_or_id, str):
        for f in FIELDS.values():
            if f.name == name_or_id:
                return f
        raise KeyError(name_or_id)
    return FIELDS[int(name_or_id)]
