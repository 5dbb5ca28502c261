"""Identities the CUDA kernels rely on to stay bit-exact while avoiding IEEE division sequences.

sgi_math.cuh computes b/255.0f and u/65535.0f as   q = x*r;  q' = fma(fma(-q, D, x), r, q)   with
r = fl(1/D).  Here every input is checked against the correctly rounded quotient using exact
rational arithmetic (no floating-point evaluation of the fma), so the claim does not depend on
any hardware."""
import math
from fractions import Fraction

import numpy as np


def round_f32(fr):
    """Fraction -> nearest-even binary32 (normal range), returned as a Python float."""
    if fr == 0:
        return 0.0
    sign = -1 if fr < 0 else 1
    fr = abs(fr)
    e = math.floor(math.log2(fr))
    while Fraction(2) ** e > fr:
        e -= 1
    while Fraction(2) ** (e + 1) <= fr:
        e += 1
    scale = Fraction(2) ** (e - 23)
    m = fr / scale
    lo = m.numerator // m.denominator
    rem = m - lo
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and lo % 2 == 1):
        lo += 1
    return sign * float(lo * scale)


def _check(D, n):
    r = Fraction(float(np.float32(1.0) / np.float32(D)))
    for b in range(n):
        q = Fraction(round_f32(Fraction(b) * r))
        rem = Fraction(round_f32(Fraction(b) - q * D))        # fma(-q, D, x): one rounding
        q2 = round_f32(rem * r + q)                            # fma(rem, r, q): one rounding
        assert q2 == float(np.float32(b) / np.float32(D)), (D, b)


def test_unorm8_identity_is_exact_for_all_256_inputs():
    _check(255, 256)


def test_unorm16_identity_is_exact_for_all_65536_inputs():
    _check(65535, 65536)


def test_plain_reciprocal_multiply_would_not_be_exact():
    # why the correction step exists
    b = np.arange(256, dtype=np.float32)
    assert np.any(b * (np.float32(1.0) / np.float32(255.0)) != b / np.float32(255.0))
