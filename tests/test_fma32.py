"""oracle/fma32.py is an exact float32 fma: checked against rational arithmetic and against fmaf."""
import ctypes
from fractions import Fraction

import numpy as np

from oracle.fma32 import fma32
from tests.util import same_bits


def _exact(a, b, c):
    """Correctly rounded float32 of a*b + c via exact rationals (float64 -> float32 of a Fraction is not
    available, so round by comparing with the two neighbouring float32 values)."""
    out = np.empty(len(a), np.float32)
    for i in range(len(a)):
        v = Fraction(float(a[i])) * Fraction(float(b[i])) + Fraction(float(c[i]))
        g = np.float32(float(v))                       # within 1 ulp; refine
        cands = [np.nextafter(g, np.float32(-np.inf)), g, np.nextafter(g, np.float32(np.inf))]
        best = None
        for x in cands:
            dist = abs(Fraction(float(x)) - v)
            even = (int(np.float32(x).view(np.uint32)) & 1) == 0
            key = (dist, 0 if even else 1)
            if best is None or key < best[0]:
                best = (key, x)
        out[i] = best[1]
    return out


def test_random_and_cancelling_inputs_match_exact_arithmetic():
    rng = np.random.RandomState(0)
    n = 4000
    a = (rng.standard_normal(n) * 10 ** rng.uniform(-3, 3, n)).astype(np.float32)
    b = (rng.standard_normal(n) * 10 ** rng.uniform(-3, 3, n)).astype(np.float32)
    c = (rng.standard_normal(n) * 10 ** rng.uniform(-3, 3, n)).astype(np.float32)
    c[:1000] = (-(a[:1000].astype(np.float64) * b[:1000].astype(np.float64))).astype(np.float32)   # heavy cancellation
    assert same_bits(fma32(a, b, c), _exact(a, b, c))


def test_constructed_double_rounding_ties():
    # p + c lands (in float64) exactly on a float32 midpoint while the exact sum is slightly off it
    cases = []
    for k in range(200):
        base = np.float32(1.0 + k * 2.0 ** -20)
        half_ulp = 2.0 ** -24
        for tiny in (2.0 ** -80, -2.0 ** -80, 2.0 ** -60, -2.0 ** -60):
            # a*b = base + half_ulp exactly needs 25 bits: use a = base + ..., choose a*b exactly representable
            a = np.float32(1.0)
            cases.append((np.float32(base), a, np.float32(tiny)))     # trivial: no tie, sanity
    a = np.array([1.0 + 2.0 ** -12] * 4, np.float32)
    b = np.array([1.0 + 2.0 ** -12] * 4, np.float32)                   # a*b = 1 + 2^-11 + 2^-24: a tie at the float32 grid near 1
    c = np.array([2.0 ** -70, -2.0 ** -70, 0.0, 2.0 ** -30], np.float32)
    got = fma32(a, b, c)
    want = _exact(a, b, c)
    assert same_bits(got, want)
    assert got[0] != got[1]                                            # the tiny addend decides the tie
    aa = np.array([x[0] for x in cases], np.float32)
    bb = np.array([x[1] for x in cases], np.float32)
    cc = np.array([x[2] for x in cases], np.float32)
    assert same_bits(fma32(aa, bb, cc), _exact(aa, bb, cc))


def test_matches_libm_fmaf(host_twin):
    rng = np.random.RandomState(1)
    n = 200000
    a = (rng.standard_normal(n) * 10 ** rng.uniform(-4, 4, n)).astype(np.float32)
    b = (rng.standard_normal(n) * 10 ** rng.uniform(-4, 4, n)).astype(np.float32)
    c = (rng.standard_normal(n) * 10 ** rng.uniform(-4, 4, n)).astype(np.float32)
    c[::3] = (-(a[::3].astype(np.float64) * b[::3].astype(np.float64))).astype(np.float32)
    a[::7] = np.float32(1.0 + 2.0 ** -12)
    b[::7] = np.float32(1.0 + 2.0 ** -12)
    c[::7] = (rng.standard_normal(len(c[::7])) * 1e-20).astype(np.float32)
    out = np.empty(n, np.float32)
    host_twin.twin_fma(ctypes.c_long(n), a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
                       c.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    assert same_bits(fma32(a, b, c), out)
    special = np.array([np.inf, -np.inf, np.nan, 0.0, -0.0, 1e38, 1e-45], np.float32)
    A, B, C = np.meshgrid(special, special, special)
    A, B, C = A.ravel().copy(), B.ravel().copy(), C.ravel().copy()
    out = np.empty(A.size, np.float32)
    host_twin.twin_fma(ctypes.c_long(A.size), A.ctypes.data_as(ctypes.c_void_p), B.ctypes.data_as(ctypes.c_void_p),
                       C.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    got = fma32(A, B, C)
    both_nan = np.isnan(got) & np.isnan(out)
    assert ((got.view(np.uint32) == out.view(np.uint32)) | both_nan).all()
