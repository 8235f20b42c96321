"""Poseidon2 width-16 / width-24 parameters for KoalaBear / BabyBear.

The round constants the reference imports from p3 (`KOALABEAR_POSEIDON2_RC_16_*`, `BABYBEAR_POSEIDON2_RC_16_*`,
/root/reference poseidon2-circuit-air/src/public_types.rs:48-53,99-104,220-226,272-278) are not in the reference tree.
They are regenerated here with the Grain LFSR of the Poseidon/Poseidon2 papers (parameters: prime field, x^alpha S-box flag 0,
n = 31 bits, t = 16, R_F = 8, R_P = 13 BabyBear / 20 KoalaBear; first R_F/2*t external, then R_P internal, then R_F/2*t external).
tests/test_oracle_kat.py pins the output against the p3 tables as recalled (BabyBear: first external row and all 13 internal
constants; KoalaBear: first 8 external constants), which match exactly.

Internal diagonal V (s_i <- V_i*s_i + sum) and rounds/S-box degree follow circuit/src/ops/poseidon2_perm/config.rs:67-75,114-122
and SURVEY.md §8c ([P3-EXT], recalled).
"""
from __future__ import annotations

import numpy as np

from .field import BABYBEAR, KOALABEAR, get_field


def grain_constants(p: int, n_bits: int, t: int, rf: int, rp: int):
    def bits(v, w):
        return [(v >> (w - 1 - i)) & 1 for i in range(w)]

    s = bits(1, 2) + bits(0, 4) + bits(n_bits, 12) + bits(t, 12) + bits(rf, 10) + bits(rp, 10) + [1] * 30

    def step():
        nb = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(nb)
        return nb

    for _ in range(160):
        step()

    def next_bit():
        while True:
            b = step()
            while b == 0:
                step()
                b = step()
            return step()

    def rnd():
        v = 0
        for _ in range(n_bits):
            v = (v << 1) | next_bit()
        return v

    out = []
    for _ in range(rf * t + rp):
        v = rnd()
        while v >= p:
            v = rnd()
        out.append(v)
    return out


# Internal diagonals as (sign, small integer, k) with k > 0 meaning 2^-k. Width 16: SURVEY.md §8c; width 24: p3-koala-bear /
# p3-baby-bear `poseidon2.rs` ([P3-EXT], recalled — the constants are injected through p3r_poseidon2_consts either way).
_DIAG = {
    (KOALABEAR, 16): [(-1, 2, 0), (1, 1, 0), (1, 2, 0), (1, 0, 1), (1, 3, 0), (1, 4, 0), (-1, 0, 1), (-1, 3, 0), (-1, 4, 0),
                      (1, 0, 8), (1, 0, 3), (1, 0, 24), (-1, 0, 8), (-1, 0, 3), (-1, 0, 4), (-1, 0, 24)],
    (BABYBEAR, 16): [(-1, 2, 0), (1, 1, 0), (1, 2, 0), (1, 0, 1), (1, 3, 0), (1, 4, 0), (-1, 0, 1), (-1, 3, 0), (-1, 4, 0),
                     (1, 0, 8), (1, 0, 2), (1, 0, 3), (1, 0, 27), (-1, 0, 8), (-1, 0, 4), (-1, 0, 27)],
    (KOALABEAR, 24): [(-1, 2, 0), (1, 1, 0), (1, 2, 0), (1, 0, 1), (1, 3, 0), (1, 4, 0), (-1, 0, 1), (-1, 3, 0), (-1, 4, 0),
                      (1, 0, 8), (1, 0, 2), (1, 0, 3), (1, 0, 4), (1, 0, 5), (1, 0, 6), (1, 0, 24),
                      (-1, 0, 8), (-1, 0, 3), (-1, 0, 4), (-1, 0, 5), (-1, 0, 6), (-1, 0, 7), (-1, 0, 9), (-1, 0, 24)],
    (BABYBEAR, 24): [(-1, 2, 0), (1, 1, 0), (1, 2, 0), (1, 0, 1), (1, 3, 0), (1, 4, 0), (-1, 0, 1), (-1, 3, 0), (-1, 4, 0),
                     (1, 0, 8), (1, 0, 2), (1, 0, 3), (1, 0, 4), (1, 0, 7), (1, 0, 9), (1, 0, 27),
                     (-1, 0, 8), (-1, 0, 2), (-1, 0, 3), (-1, 0, 4), (-1, 0, 5), (-1, 0, 6), (-1, 0, 7), (-1, 0, 27)],
}
# partial rounds (circuit/src/ops/poseidon2_perm/config.rs:67-122): width 16: 13 / 20, width 24: 21 / 23
_ROUNDS_P = {(BABYBEAR, 16): 13, (KOALABEAR, 16): 20, (BABYBEAR, 24): 21, (KOALABEAR, 24): 23}


class Poseidon2Params:
    def __init__(self, field_id, width: int = 16):
        f = get_field(field_id)
        self.field = f
        p = f.p
        if (f.field_id, width) not in _DIAG:
            raise ValueError((field_id, width))
        self.width = width
        self.rounds_f = 8
        self.sbox_degree = 3 if f.field_id == KOALABEAR else 7
        self.rounds_p = _ROUNDS_P[(f.field_id, width)]
        diag = _DIAG[(f.field_id, width)]
        rc = grain_constants(p, 31, width, self.rounds_f, self.rounds_p)
        half = self.rounds_f // 2 * width
        self.external_rc = np.array(rc[:half] + rc[half + self.rounds_p:], dtype=np.uint32)  # initial then terminal
        self.internal_rc = np.array(rc[half:half + self.rounds_p], dtype=np.uint32)
        d = []
        for sign, num, sh in diag:  # (sign, small integer, k) with k>0 meaning 1/2^k
            v = num % p if sh == 0 else pow(pow(2, sh, p), p - 2, p)
            d.append(v if sign > 0 else (p - v) % p)
        self.internal_diag = np.array(d, dtype=np.uint32)

    # vectorised reference permutation on canonical uint64 arrays of shape (n, width); returns all intermediate
    # states when `trace=True` (used by the Poseidon2 table trace generator).
    def _sbox(self, x):
        p = np.uint64(self.field.p)
        x2 = x * x % p
        if self.sbox_degree == 3:
            return x2 * x % p
        x3 = x2 * x % p
        x4 = x2 * x2 % p
        return x3 * x4 % p

    def _external(self, s):
        p = np.uint64(self.field.p)
        s = s.reshape(-1, self.width // 4, 4)
        a, b, c, d = s[:, :, 0], s[:, :, 1], s[:, :, 2], s[:, :, 3]
        o = np.stack([2 * a + 3 * b + c + d, a + 2 * b + 3 * c + d, a + b + 2 * c + 3 * d, 3 * a + b + c + 2 * d], axis=2) % p
        sums = o.sum(axis=1) % p
        o = (o + sums[:, None, :]) % p
        return o.reshape(-1, self.width)

    def permute(self, states):
        p = np.uint64(self.field.p)
        w = self.width
        s = np.asarray(states, dtype=np.uint64).reshape(-1, w) % p
        s = self._external(s)
        half = self.rounds_f // 2
        erc = self.external_rc.astype(np.uint64).reshape(self.rounds_f, w)
        for r in range(half):
            s = self._external(self._sbox((s + erc[r]) % p))
        diag = self.internal_diag.astype(np.uint64)
        for r in range(self.rounds_p):
            s[:, 0] = self._sbox((s[:, 0] + np.uint64(self.internal_rc[r])) % p)
            tot = s.sum(axis=1) % p
            s = (tot[:, None] + diag * s % p) % p
        for r in range(half, self.rounds_f):
            s = self._external(self._sbox((s + erc[r]) % p))
        return s.astype(np.uint32)

    def sponge(self, rows, rate: int, out: int = 8):
        """PaddingFreeSponge<Perm, width, rate, out> in overwrite mode over the rows of a canonical (n, c) matrix."""
        rows = np.asarray(rows, dtype=np.uint64)
        st = np.zeros((rows.shape[0], self.width), dtype=np.uint64)
        for c0 in range(0, rows.shape[1], rate):
            chunk = rows[:, c0:c0 + rate]
            st[:, :chunk.shape[1]] = chunk
            st = self.permute(st).astype(np.uint64)
        return st[:, :out].astype(np.uint32)
