"""CPU model of the proofs the three-warp forward sweep relies on (`duo::prove` in csrc/qr_banded.cu): a candidate square root /
quotient is accepted iff an exact residual test passes.  The device operations are restated here one by one — `fma` as exact
rational arithmetic followed by ONE rounding, the threshold's exponent arithmetic on the bit pattern, the margins' high words —
and the verdict is compared with the ground truth from exact integer / rational arithmetic: the test must accept the correctly
rounded value and reject every neighbour, including candidates that are powers of two (asymmetric neighbourhood), results that
sit exactly on the acceptance boundary (s = 1 + 2^-52, sqrt = 1), and garbage (zero, negative, Inf, NaN)."""
import math
import struct
from fractions import Fraction

import numpy as np
import pytest


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def from_bits(b):
    return struct.unpack("<d", struct.pack("<Q", b & 0xFFFFFFFFFFFFFFFF))[0]


def rn(fr):
    """Fraction -> nearest double (ties to even): CPython's Fraction.__float__ rounds correctly."""
    return float(fr)


def fma(a, b, c):
    if not all(map(math.isfinite, (a, b, c))):
        return a * b + c
    return rn(Fraction(a) * Fraction(b) + Fraction(c))


def ulp_pow2(x, lower, sh, mask):
    """duo::ulp_pow2<SH, MASK>: 2^(e - SH) for x in [2^e, 2^(e+1)), one binade less when x is a power of two and `lower` is set."""
    b = bits(x)
    hi, lo = b >> 32, b & 0xFFFFFFFF
    if mask:
        hi &= 0x7FFFFFFF
    t = lo - lower
    borrow = 1 if t < 0 else 0
    h2 = (hi - (sh << 20) - borrow) & 0xFFFFFFFF
    return from_bits((h2 & 0xFFF00000) << 32)


def hi_word(x):
    return bits(x) >> 32


def margin_ok(ms):
    h = max(hi_word(m) for m in ms)
    return h < 0x7FF00000


def prove_sqrt(s, n):
    d2 = fma(-n, n, s)
    tn = n * ulp_pow2(n, hi_word(d2) >> 31, 52, False)
    m = fma(d2, 2.0 ** -60, tn - abs(d2))
    sbox = (hi_word(s) - (123 << 20)) & 0xFFFFFFFF
    return margin_ok([n, m]) and sbox <= (1800 << 20)   # (the candidate itself: positive and finite)


def prove_div(p, q, x):
    """x == RN(p / q), q > 0 (the kernel's X = |a| + n and n are positive)."""
    r = fma(x, -q, p)
    lower = ((hi_word(r) ^ hi_word(x)) >> 31) & 1
    t = q * ulp_pow2(x, lower, 53, True)
    return margin_ok([t - abs(r)])


def true_sqrt(s):
    """correctly rounded sqrt of a double by exact integer arithmetic"""
    fr = Fraction(s)
    # scale to an integer with plenty of bits
    k = 240
    v = fr * (1 << (2 * k))
    r = math.isqrt(v.numerator // v.denominator)
    cand = float(Fraction(r, 1 << k))
    best = None
    for c in (np.nextafter(cand, -np.inf), cand, np.nextafter(cand, np.inf)):
        lo_, hi_ = Fraction(float(c)) - Fraction(float(c) - float(np.nextafter(c, -np.inf))) / 2, Fraction(float(c)) + Fraction(float(np.nextafter(c, np.inf)) - float(c)) / 2
        if lo_ * lo_ < fr < hi_ * hi_ or (lo_ * lo_ == fr) or (hi_ * hi_ == fr):
            best = float(c)
    assert best is not None
    return best


def neighbours(x, k=2):
    out = [x]
    a = b = x
    for _ in range(k):
        a = float(np.nextafter(a, -np.inf))
        b = float(np.nextafter(b, np.inf))
        out += [a, b]
    return out


def test_sqrt_proof_accepts_exactly_the_correctly_rounded_root():
    rng = np.random.default_rng(1)
    radicands = list(np.exp(rng.uniform(-60, 60, size=600)))
    radicands += [1.0, 4.0, 0.25, 1.0 + 2.0 ** -52, 1.0 + 2.0 ** -51, 1.0 - 2.0 ** -53, 4.0 - 2.0 ** -51, 2.0, 3.0, 1e-30, 1e30,
                  float(np.nextafter(16.0, 0)), float(np.nextafter(16.0, 100)), 1.0 + 3 * 2.0 ** -52]
    # radicands of the form a^2 + 1 with tiny a (the Bessel sweep at large z)
    radicands += [float(a * a + 1.0) for a in (1e-9, 1.05e-8, 1.5e-8, 2e-8, 3e-8, 1e-7)]
    for s in radicands:
        s = float(s)
        n0 = true_sqrt(s)
        assert n0 == float(np.sqrt(s))
        for n in neighbours(n0):
            assert prove_sqrt(s, n) == (n == n0), (s, n, n0)
    for bad in (0.0, -1.0, float("inf"), float("nan"), 1e-320):
        assert not prove_sqrt(2.0, bad)
        assert not prove_sqrt(2.0, -np.sqrt(2.0))


def test_div_proof_accepts_exactly_the_correctly_rounded_quotient():
    rng = np.random.default_rng(2)
    cases = [(float(p), float(q)) for p, q in zip(rng.normal(size=500) * np.exp(rng.uniform(-40, 40, size=500)), np.exp(rng.uniform(-30, 30, size=500)))]
    # quotients that are (or round to) powers of two, from both sides; tau = X / n in [1, 2]
    cases += [(1.0, 1.0), (2.0, 1.0), (1.0, 2.0), (float(np.nextafter(1.0, 2)), 1.0), (float(np.nextafter(1.0, 0)), 1.0), (1.0, float(np.nextafter(1.0, 2))),
              (1.0, float(np.nextafter(1.0, 0))), (-1.0, float(np.nextafter(2.0, 0))), (3.0, 3.0), (-7.5, 2.5), (1.0 + 2.0 ** -52, 1.0 + 2.0 ** -51),
              (2.0 - 2.0 ** -52, 1.0), (1.0 + 2.0 ** -30, 1.0 + 2.0 ** -31)]
    for p, q in cases:
        x0 = rn(Fraction(p) / Fraction(q))
        assert x0 == p / q
        for x in neighbours(x0):
            assert prove_div(p, q, x) == (x == x0), (p, q, x, x0)
        assert not prove_div(p, q, -x0) or x0 == 0.0
    for bad in (0.0, float("inf"), float("nan"), 1e-320):
        assert not prove_div(1.5, 2.5, bad)


def test_threshold_is_an_exact_power_of_two_scaling():
    for x in (1.0, 1.5, float(np.nextafter(2.0, 0)), 3.7e-20, 8.1e30):
        e = math.frexp(x)[1] - 1
        assert ulp_pow2(x, 0, 52, False) == 2.0 ** (e - 52)
        assert ulp_pow2(x, 0, 53, True) == 2.0 ** (e - 53)
        assert ulp_pow2(-x, 0, 53, True) == 2.0 ** (e - 53)
        pow2 = math.frexp(x)[0] == 0.5
        assert ulp_pow2(x, 1, 52, False) == 2.0 ** (e - 52 - (1 if pow2 else 0))
    # zero, subnormal and negative (unmasked) candidates get a negative threshold: every margin test fails
    assert ulp_pow2(0.0, 0, 53, True) < 0 and ulp_pow2(0.0, 1, 53, True) < 0 and ulp_pow2(1e-320, 0, 52, False) < 0 and ulp_pow2(-2.0, 0, 52, False) < 0
