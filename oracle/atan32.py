"""float32 atan2 with a pinned operation order (T1 bit model of the planar controllers).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference's 2D controllers evaluate
``math.atan2(desired_acc[1], desired_acc[0])`` in float64
(/root/reference/gym_reinmav/envs/native/quadrotor2d.py:128, quadrotor2d_slungload.py:173).  A float32 CUDA path
cannot call the host libm, and the CUDA library's ``atan2f`` is not reproducible in NumPy, so the float32 model fixes
one algorithm that both sides implement with individually rounded IEEE operations in a fixed order (no fused
multiply-add anywhere): the classic single-precision scheme (Cephes ``atanf`` / ``atan2f``) — octant reduction
``x -> -1/x`` above tan(3pi/8), ``x -> (x-1)/(x+1)`` above tan(pi/8), a degree-4 polynomial in x^2, quadrant offset
0 / +-pi.  Max error 2.5e-7 absolute on [-pi, pi] (tests/test_host_twin.py).  csrc/rmv_math.cuh ``atan2_pinned``
is the same sequence.
"""
import numpy as np

F = np.float32
PI, PIO2, PIO4 = F(3.14159265358979323846), F(1.57079632679489661923), F(0.78539816339744830962)
T3P8, TP8 = F(2.414213562373095), F(0.4142135623730950)
A4, A3, A2, A1 = F(8.05374449538e-2), F(1.38776856032e-1), F(1.99777106478e-1), F(3.33329491539e-1)


def atan32(x):
    """atan(x), every operation rounded to float32."""
    x = np.asarray(x, dtype=F)
    with np.errstate(all="ignore"):
        neg = np.signbit(x)
        ax = np.abs(x)
        big, mid = ax > T3P8, ax > TP8
        y0 = np.where(big, PIO2, np.where(mid, PIO4, F(0.0)))
        r = np.where(big, -(F(1.0) / ax), np.where(mid, (ax - F(1.0)) / (ax + F(1.0)), ax))
        z = r * r
        p = (((A4 * z - A3) * z + A2) * z - A1) * z * r + r
        y = y0 + p
    return np.where(neg, -y, y).astype(F)


def atan2_32(y, x):
    """atan2(y, x), every operation rounded to float32; NaN in -> NaN out."""
    y, x = np.asarray(y, dtype=F), np.asarray(x, dtype=F)
    with np.errstate(all="ignore"):
        z = atan32(y / x)
        w = np.where(x < 0, np.where(y < 0, -PI, PI), F(0.0))
        out = w + z
        out = np.where(y == 0, np.where(x < 0, np.where(np.signbit(y), -PI, PI), y), out)    # +-0 -> +-0 or +-pi, as libm
        out = np.where(x == 0, np.where(y < 0, -PIO2, np.where(y == 0, F(0.0), PIO2)), out)
        out = np.where(np.isnan(x) | np.isnan(y), F(np.nan), out)
    return out.astype(F)
