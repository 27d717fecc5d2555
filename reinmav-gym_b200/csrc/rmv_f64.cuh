// rmv_f64.cuh — float64 asin / atan2 sized for the FP64 pipe of k_reinmav_step (rmv_reinmav.cuh).
//
// reinmav-v0's controller reads three Euler angles per Euler substep (RotToRPY, gym_reinmav/envs/native/
// reinmav_env.py:341-346): one asin and two atan2, 50 times per env-step.  The CUDA math library's versions cost
// ~230 of the ~370 FP64 instructions of a substep (ncu, profiles/r2_ncu_reinmav_n20.txt: the FP64 pipe is the
// kernel's limiter), mostly IEEE divisions and wide argument ranges this path does not need.  These versions keep
// double accuracy (a few 1e-16 absolute, tests/test_host_twin.py against libm) with about 65:
//   atan2(y, x)  octant reduction on |x|, |y| plus the pi/8 split  atan(m/M) = pi/4 + atan((m - M)/(m + M)),  both
//                applied BEFORE the one division, so the polynomial only sees |t| <= tan(pi/8): 10 terms in t^2
//   asin(x)      |x| < 1/2: x + x^3 P(x^2), 12 terms; beyond (roll above 30 degrees) the library's asin
//   the division MUFU.RCP64H seed, one Newton step, one residual correction (5 FP64 ops, no slow path)
// Coefficients: Chebyshev-node interpolants of (f(sqrt u)/sqrt u - 1)/u computed at 50 digits, rounded to double.
// Host build (tests/host_twin): same code with an exact division.
#pragma once
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define RMV_F64_HD __host__ __device__ __forceinline__
#else
#define RMV_F64_HD inline
#endif

namespace rmv {

// the rare paths, kept out of line so the substep loop stays short
#if defined(__CUDACC__)
__device__ __noinline__ double atan2_library(double y, double x) { return ::atan2(y, x); }
__device__ __noinline__ double asin_library(double x) { return ::asin(x); }
__device__ __noinline__ double div_library(double n, double d) { return n / d; }
#endif
RMV_F64_HD double div_rare(double n, double d) {
#if defined(__CUDA_ARCH__)
  return div_library(n, d);
#else
  return n / d;
#endif
}
RMV_F64_HD double atan2_rare(double y, double x) {
#if defined(__CUDA_ARCH__)
  return atan2_library(y, x);
#else
  return ::atan2(y, x);
#endif
}
RMV_F64_HD double asin_rare(double x) {
#if defined(__CUDA_ARCH__)
  return asin_library(x);
#else
  return ::asin(x);
#endif
}

// 2^-930 <= |x| < 2^931, decided on the exponent field (an integer compare: the FP64 pipe is the scarce one)
RMV_F64_HD bool normal_range(double x) {
#if defined(__CUDA_ARCH__)
  const unsigned hi = (unsigned)__double2hiint(x);
#else
  unsigned long long bits;
  memcpy(&bits, &x, sizeof(bits));
  const unsigned hi = (unsigned)(bits >> 32);
#endif
  return ((hi >> 20) & 0x7ffu) - 93u < 1861u;
}

// n / d for a normal, positive d
RMV_F64_HD double div_newton(double n, double d) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = ::fma(r, ::fma(-d, r, 1.0), r);              // seed is good to 2^-19.9 (tools/f64_check.cu): 2^-39.8 after one Newton step
  const double q = n * r;
  return ::fma(::fma(-d, q, n), r, q);             // the residual correction squares that again: within one ulp of n / d
#else
  return n / d;
#endif
}

// Coefficient tables live in the constant bank on the device: a DFMA takes a c[bank][offset] operand directly, while a
// 64-bit literal costs two uniform-register moves per use.
#if defined(__CUDA_ARCH__)
#define RMV_F64_TABLE __constant__
#else
#define RMV_F64_TABLE static const
#endif
RMV_F64_TABLE double kAtanCoef[10] = {-0x1.5555555555546p-2, 0x1.9999999990a9fp-3, -0x1.2492491dd4b29p-3, 0x1.c71c6dd1ca28cp-4,
                                      -0x1.745c7f8421f5cp-4, 0x1.3b068efc532d3p-4, -0x1.105e40d85ead1p-4, 0x1.d5eb9ad0804d0p-5,
                                      -0x1.6f4657797fec4p-5, 0x1.74bea059cee82p-6};
RMV_F64_TABLE double kAsinCoef[12] = {0x1.555555555554fp-3, 0x1.3333333336da5p-4, 0x1.6db6db684b6a1p-5, 0x1.f1c71f95269afp-6,
                                      0x1.6e8b2b3b10be4p-6, 0x1.1c593c7b1d958p-6, 0x1.c87265d47ef49p-7, 0x1.8522ddffa6208p-7,
                                      0x1.ff5fc4d14c735p-8, 0x1.06b9d26d10838p-6, -0x1.603991d6060e0p-7, 0x1.cd864394d2ff2p-6};

// atan(t) for |t| <= tan(pi/8):  t + t^3 P(t^2)
RMV_F64_HD double atan_small(double t) {
  const double u = t * t;
  double p = kAtanCoef[9];
#pragma unroll
  for (int k = 8; k >= 0; --k) p = ::fma(p, u, kAtanCoef[k]);
  return ::fma(t * u, p, t);
}

// a > b for two doubles with clear sign bits: their bit patterns order like integers
RMV_F64_HD bool gt_nonneg(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __double_as_longlong(a) > __double_as_longlong(b);
#else
  long long ia, ib;
  memcpy(&ia, &a, sizeof(ia));
  memcpy(&ib, &b, sizeof(ib));
  return ia > ib;
#endif
}

// min(max(v, 0), hi) for hi > 0, decided on the bit patterns (NaN: hi or 0 by its sign bit)
RMV_F64_HD double clamp_zero_to(double v, double hi) {
  if (::signbit(v)) return 0.0;
  return gt_nonneg(v, hi) ? hi : v;
}

// pi/4 * {0, 1, 2, 1, 4, 3, 2, 3}: where the reduced angle is measured from, by (x < 0, |y| > |x|, beyond pi/8)
RMV_F64_TABLE double kOctantBase[8] = {0.0, 0x1.921fb54442d18p-1, 0x1.921fb54442d18p+0, 0x1.921fb54442d18p-1,
                                       0x1.921fb54442d18p+1, 0x1.2d97c7f3321d2p+1, 0x1.921fb54442d18p+0, 0x1.2d97c7f3321d2p+1};

// atan2 for max(|x|, |y|) in the normal range
RMV_F64_HD bool atan2_fast_ok(double y, double x) {
  const double ax = ::fabs(x), ay = ::fabs(y);
  return normal_range(gt_nonneg(ay, ax) ? ay : ax);                // false for zeros, subnormals, inf, NaN (NaN orders above inf)
}
RMV_F64_TABLE double kTanPi8 = 0x1.a827999fcef32p-2;
RMV_F64_HD double atan2_core(double y, double x) {
  const double ax = ::fabs(x), ay = ::fabs(y);
  const bool swap = gt_nonneg(ay, ax);
  const double big = swap ? ay : ax, small = swap ? ax : ay;
  const bool upper = gt_nonneg(small, kTanPi8 * big);  // beyond pi/8: measure from pi/4
  const bool xneg = ::signbit(x);
  const double num = upper ? small - big : small;
  const double den = upper ? small + big : big;
  // |angle| = base + sign * atan(num / den); the mirror images (about pi/4 for swap, about pi/2 for x < 0) flip the sign
  const double t = div_newton(num, den);
  const double a = atan_small(swap != xneg ? -t : t) + kOctantBase[(xneg ? 4 : 0) + (swap ? 2 : 0) + (upper ? 1 : 0)];
  return ::copysign(a, y);
}
RMV_F64_HD double atan2_f64(double y, double x) {
  return atan2_fast_ok(y, x) ? atan2_core(y, x) : atan2_rare(y, x);   // rare: the library's rules
}

// asin for |x| < 1/2
RMV_F64_HD bool asin_fast_ok(double x) { return gt_nonneg(0.5, ::fabs(x)); }   // false for NaN too
RMV_F64_HD double asin_core(double x) {
  const double u = x * x;
  double p = kAsinCoef[11];
#pragma unroll
  for (int k = 10; k >= 0; --k) p = ::fma(p, u, kAsinCoef[k]);
  return ::fma(x * u, p, x);
}
RMV_F64_HD double asin_f64(double x) { return asin_fast_ok(x) ? asin_core(x) : asin_rare(x); }

}  // namespace rmv
