"""Exact float32 fused multiply-add in NumPy (TEST INFRASTRUCTURE, see oracle/__init__.py).

The CUDA path fuses ``a*b + c`` into one FFMA where the arithmetic header says so
(csrc/rmv_math.cuh ``fma``).  NumPy has no fma, so the float32 model evaluates it exactly:

  * ``p = a*b`` is exact in float64 (24 x 24 significant bits),
  * ``s = fl64(p + c)`` with its exact rounding error ``err`` from Knuth's TwoSum,
  * ``float32(s)`` equals the correctly rounded ``p + c`` unless ``s`` sits exactly on a float32
    rounding boundary (midpoint of two adjacent float32 values) while ``err != 0``; then the tie is
    broken by the sign of ``err`` instead of round-to-even.

tests/test_fma32.py checks it against exact rational arithmetic, including constructed midpoint cases,
and against glibc ``fmaf`` through the host build of the kernel header.
"""
import numpy as np

F32, F64 = np.float32, np.float64


def fma32(a, b, c):
    a, b, c = np.broadcast_arrays(np.asarray(a, F32), np.asarray(b, F32), np.asarray(c, F32))
    a64, b64, c64 = a.astype(F64), b.astype(F64), c.astype(F64)
    with np.errstate(all="ignore"):
        p = a64 * b64
        s = p + c64
        bb = s - p
        err = (p - (s - bb)) + (c64 - bb)
        r = s.astype(F32)
        # candidates: inexact float64 sum whose float32 rounding was a tie
        cand = (err != 0) & np.isfinite(s) & np.isfinite(r)
        if cand.any():
            bits = np.ascontiguousarray(s).view(np.uint64)
            normal = np.abs(s) >= F64(2.0 ** -126)
            tie_pattern = (bits & np.uint64(0x1FFFFFFF)) == np.uint64(0x10000000)
            cand &= np.where(normal, tie_pattern, True)          # float32-subnormal results: always take the slow path
        if cand.any():
            idx = np.nonzero(cand)
            si, ei, ri = s[idx], err[idx], r[idx]
            r64 = ri.astype(F64)
            d = si - r64
            toward = np.where(d > 0, F32(np.inf), F32(-np.inf)).astype(F32)
            nb = np.nextafter(ri, toward)
            mid = (r64 + nb.astype(F64)) * 0.5
            fix = (d != 0) & (si == mid) & (np.sign(ei) == np.sign(d))
            # a tie that round-to-even resolved AWAY from the true value the other way round:
            # s == midpoint below/above r but r was chosen; if err points back towards r nothing changes
            ri = np.where(fix, nb, ri)
            # exact float32 value with an error term: s == r exactly, err cannot move it
            r = r.copy()
            r[idx] = ri
    return r
