"""float32 sine/cosine with a pinned operation order (T1 bit model).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference evaluates
``math.cos(att + pi/2)`` / ``math.sin(att + pi/2)`` in float64
(/root/reference/gym_reinmav/envs/native/quadrotor2d.py:88,
quadrotor2d_slungload.py:97,108,129).  A float32 CUDA path cannot call the host
libm, so the float32 model fixes one algorithm that both sides implement with
IEEE operations in a fixed order — products, and fused multiply-adds where the
kernel issues an FFMA (oracle/fma32.py is an exact float32 fma): Cody-Waite
reduction by pi/2 in three parts followed by the classic single-precision
minimax polynomials (Cephes sinf/cosf coefficients).  Max error ~1.5 ulp for |x| <= 1e5; larger
arguments reduce in float64.  csrc/rmv_math.cuh ``rmv_sincos`` is the same
sequence.
"""
import numpy as np

from .fma32 import fma32

F = np.float32
TWO_OVER_PI = F(0.63661977236758134)
P1 = F(1.5703125)                    # pi/2 head, 8 significant bits: k*P1 exact for |k| < 2^16
P2 = F(4.837512969970703125e-4)
P3 = F(7.54978995489188216e-8)
S1, S2, S3 = F(-1.6666654611e-1), F(8.3321608736e-3), F(-1.9515295891e-4)
C1, C2, C3 = F(4.166664568298827e-2), F(-1.388731625493765e-3), F(2.443315711809948e-5)
BIG = F(1.0e5)
PIO2_HI = np.float64(1.5707963267948966)
PIO2_LO = np.float64(6.123233995736766e-17)
TWO_OVER_PI_D = np.float64(0.6366197723675814)


def sincos32(x):
    """Return (sin x, cos x) as float32 arrays; every op rounded to float32."""
    x = np.asarray(x, dtype=F)
    with np.errstate(invalid="ignore", over="ignore"):
        kf = np.rint(x * TWO_OVER_PI)
        r = fma32(-kf, P3, fma32(-kf, P2, fma32(-kf, P1, x)))
        quad = kf.astype(np.int64) & 3 if np.all(np.isfinite(kf)) else \
            np.where(np.isfinite(kf), kf, F(0)).astype(np.int64) & 3
        big = ~(np.abs(x) <= BIG)                 # also catches NaN
        if np.any(big):
            xd = x.astype(np.float64)
            kd = np.rint(xd * TWO_OVER_PI_D)
            rd = (xd - kd * PIO2_HI) - kd * PIO2_LO
            qd = kd - np.float64(4.0) * np.floor(kd * np.float64(0.25))
            qd = np.where(np.isfinite(qd), qd, 0.0).astype(np.int64)
            r = np.where(big, rd.astype(F), r)
            quad = np.where(big, qd, quad)
        z = r * r
        t = fma32(S3, z, S2)
        t = fma32(t, z, S1)
        t = t * z
        sr = fma32(t, r, r)
        t = fma32(C3, z, C2)
        t = fma32(t, z, C1)
        t = t * z
        t = t * z
        t = fma32(F(-0.5), z, t)
        cr = t + F(1.0)
    s = np.where(quad == 0, sr, np.where(quad == 1, cr, np.where(quad == 2, -sr, -cr)))
    c = np.where(quad == 0, cr, np.where(quad == 1, -sr, np.where(quad == 2, -cr, sr)))
    return s.astype(F), c.astype(F)
