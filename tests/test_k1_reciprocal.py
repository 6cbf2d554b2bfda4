"""K1 computes 1/(raw/16 + 1e-6) with an integer seed and four Newton steps in
FMAs (no MUFU, csrc/k1_disp2depth.cu rcp_rn_fma).  The input domain is finite,
so the claim "equals IEEE float32 division for every input" is checked here
exhaustively with exact rational arithmetic (CPU only)."""
import math
import struct
from fractions import Fraction

import numpy as np


def _round_f32(fr):
    if fr == 0:
        return 0.0
    sign = -1 if fr < 0 else 1
    a = abs(fr)
    e = math.floor(math.log2(a.numerator) - math.log2(a.denominator))
    while Fraction(2) ** e > a:
        e -= 1
    while Fraction(2) ** (e + 1) <= a:
        e += 1
    scaled = a / Fraction(2) ** (e - 23)
    n = scaled.numerator // scaled.denominator
    rem = scaled - n
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (n & 1)):
        n += 1
    return sign * float(Fraction(n) * Fraction(2) ** (e - 23))


def _fma(a, b, c):
    return _round_f32(Fraction(a) * Fraction(b) + Fraction(c))


def _u2f(u):
    return struct.unpack('<f', struct.pack('<I', u & 0xffffffff))[0]


def _f2u(f):
    return struct.unpack('<I', struct.pack('<f', f))[0]


def test_fma_reciprocal_is_correctly_rounded_for_every_raw_code():
    raws = np.arange(0, 65535, dtype=np.float32)
    xs = raws * np.float32(0.0625) + np.float32(1e-6)
    ref = np.float32(1.0) / xs
    for x, want in zip(xs.tolist() + [1e-6], ref.tolist() + [float(np.float32(1) / np.float32(1e-6))]):
        x = float(np.float32(x))
        r = _u2f(0x7EF311C7 - _f2u(x))
        for _ in range(4):
            e = _fma(-x, r, 1.0)
            r = _fma(r, e, r)
        assert r == want, (x, r, want)


def test_u16_to_float_mantissa_trick_is_exact():
    raw = np.arange(0, 65536, dtype=np.uint32)
    bits = (np.uint32(0x4B000000) | raw).view(np.float32) - np.float32(8388608.0)
    np.testing.assert_array_equal(bits, raw.astype(np.float32))


def test_multiply_high_division_is_exact():
    for d in [1, 3, 80, 160, 240, 1000, 65535]:
        magic = ((1 << 40) + d - 1) // d
        n = np.concatenate([np.arange(0, 4096), np.random.default_rng(d).integers(0, 1 << 24, 20000),
                            [(1 << 24) - 1]]).astype(object)
        for v in n[:: max(1, len(n) // 3000)]:
            assert (int(v) * magic) >> 40 == int(v) // d
