"""The divider-free / table-free formulas of gandtr_b200/csrc/clahe_math.cuh restated with exact rational arithmetic and
checked against the reference's float32 expressions over their WHOLE domains:
  lab_cell_u8      (v*32768 + 255) // 510          == cvRound(float32(v)/255 * 16384)            v in [0, 255]
                   (v*514 + 4) >> 8                == cvRound(...) >> 5   (t << 4 | f)            v in [0, 255]
  lab_l8_int       (o*255) >> 14                   == trunc(((o*2^-14*100)/100) * 255)            o in [0, 16384]
  lab_l8_fast      div_by_const<1>(L, 100)         == L / 100                                     o in [0, 16384]
  lab_chroma_fast  div_by_const<1>(o/64, 255)      == (a + 128) / 255                             o in [0, 16384]
  lab_l_from_u8    div_by_const<1>(v, 255)         == float32(v) / 255                            v in [0, 255]
"""
from fractions import Fraction

import numpy as np

f32 = np.float32


def rn(fr):
    """Fraction -> nearest-even float32."""
    if fr == 0:
        return f32(0)
    x = f32(float(fr))
    cands = [x, np.nextafter(x, f32(np.inf)), np.nextafter(x, f32(-np.inf))]
    return min(cands, key=lambda v: (abs(Fraction(float(v)) - fr), int(v.view(np.uint32)) & 1))


def fma(a, b, c):
    return rn(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def div_by_const(a, b, iters):
    r = f32(1) / f32(b)
    q = f32(a) * r
    for _ in range(iters):
        e = fma(-f32(b), q, a)
        q = fma(e, r, q)
    return q


def test_u8_cell_quantisation_integer_formula():
    for v in range(256):
        ref = int(np.rint((f32(v) / f32(255)) * f32(16384)))
        assert (v * 32768 + 255) // 510 == ref
        assert (v * 514 + 4) >> 8 == ref >> 5          # lattice index t = ref >> 9 (0..32), fraction f = (ref >> 5) & 15


def test_lightness_byte_integer_formula():
    for o in range(0, 16385):
        L = (f32(o) * f32(1.0 / 16384.0)) * f32(100)
        assert int((L / f32(100)) * f32(255)) == (o * 255) >> 14


def test_lightness_and_chroma_divisions_exact_on_their_domains():
    for o in range(0, 16385):
        L = (f32(o) * f32(1.0 / 16384.0)) * f32(100)
        assert div_by_const(L, 100.0, 1) == L / f32(100)
        x = f32(o) * f32(1.0 / 64.0)
        a = (f32(o) * f32(1.0 / 16384.0)) * f32(256) - f32(128)
        assert a + f32(128) == x                                   # (a + 128) is exactly o / 64
        assert div_by_const(x, 255.0, 1) == x / f32(255)
    for v in range(256):
        assert div_by_const(f32(v), 255.0, 1) == f32(v) / f32(255)


def test_two_step_division_matches_ieee_on_random_operands():
    rs = np.random.RandomState(0)
    for std in (0.229, 0.224, 0.225):
        for a in (rs.rand(300).astype(np.float32) - f32(0.485)):
            assert div_by_const(a, std, 2) == a / f32(std)
