/* pm_math.h -- a small PORTABLE double-precision libm: sin, cos, tan, log, hypot built only from IEEE-754
 * correctly rounded operations (+ - * / sqrt), integer bit manipulation and rint, with a fixed evaluation order.
 *
 * Why it exists.  The reference computes sin / cos / tan / log / hypot with glibc, the CUDA kernel with CUDA's
 * libdevice; both are accurate to about an ulp but round differently, and a CILQR solve amplifies such
 * last-bit differences on a few ill-conditioned scenarios.  To separate "the kernel
 * implements the reference's logic" from "two libms round differently", the STRICT build of the kernel
 * (-DCILQR_STRICT=1 -fmad=false) and the oracle built with -DCILQR_PM_LIBM call THESE functions instead: the same
 * source, compiled by nvcc for the device and by gcc (-ffp-contract=off) for the host, gives bit-identical results
 * on both sides, so strict GPU output can be compared with that oracle bit for bit.  Accuracy: <= 2 ulp for
 * sin / cos / log / hypot and <= 3 ulp for tan on the ranges the solver uses (tests/test_pm_math.py measures it
 * against glibc) -- an ordinary libm, just a portable one.
 *
 * Not used by the production kernel or the default oracle.  C99 and CUDA C++.
 */
#ifndef CILQR_PM_MATH_H_
#define CILQR_PM_MATH_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define PM_FN __host__ __device__ static inline
#else
#define PM_FN static inline
#endif

PM_FN uint64_t pm_bits(double x) {
  uint64_t u;
  memcpy(&u, &x, sizeof(u));
  return u;
}
PM_FN double pm_from_bits(uint64_t u) {
  double x;
  memcpy(&x, &u, sizeof(x));
  return x;
}

/* sin and cos of r, |r| <= pi/4 (+ a few ulp): Taylor polynomials in z = r*r, Horner, through x^17 / x^18
 * (truncation < 1e-17 relative). */
PM_FN double pm_ksin(double r) {
  const double z = r * r;
  double p = 2.8114572543455206e-15;
  p = p * z + -7.647163731819816e-13;
  p = p * z + 1.6059043836821613e-10;
  p = p * z + -2.505210838544172e-08;
  p = p * z + 2.7557319223985893e-06;
  p = p * z + -0.0001984126984126984;
  p = p * z + 0.008333333333333333;
  p = p * z + -0.16666666666666666;
  return r + r * (z * p);
}
PM_FN double pm_kcos(double r) {
  const double z = r * r;
  double p = -1.5619206968586225e-16;
  p = p * z + 4.779477332387385e-14;
  p = p * z + -1.1470745597729725e-11;
  p = p * z + 2.08767569878681e-09;
  p = p * z + -2.755731922398589e-07;
  p = p * z + 2.48015873015873e-05;
  p = p * z + -0.001388888888888889;
  p = p * z + 0.041666666666666664;
  const double hz = 0.5 * z;
  const double w = 1.0 - hz;
  return w + (((1.0 - w) - hz) + z * (z * p));
}

/* r = x - n*pi/2 with a three-part pi/2 (33 + 33 + 53 bits), n = rint(x * 2/pi); returns n mod 4.
 * |x| >= 1e5 first goes through fmod(x, 2*pi as a double): deterministic, not accurate -- such arguments only
 * occur in rollouts that have already blown up, where the result is rejected whatever it is. */
PM_FN int pm_rem_pio2(double x, double* r) {
  const double pio2_1 = 1.5707963267341256;      /* 0x1.921fb54400000p+0  */
  const double pio2_2 = 6.077100506303966e-11;   /* 0x1.0b4611a600000p-34 */
  const double pio2_3 = 2.0222662487959506e-21;  /* 0x1.3198a2e037073p-69 */
  const double invpio2 = 0.6366197723675814;
  if (!(fabs(x) < 1e5)) x = fmod(x, 6.283185307179586);
  const double fn = rint(x * invpio2);
  double y = x - fn * pio2_1;
  y = y - fn * pio2_2;
  y = y - fn * pio2_3;
  *r = y;
  return (int)((int64_t)fn & 3);
}

PM_FN void pm_sincos(double x, double* s, double* c) {
  if (!(fabs(x) <= 1.7976931348623157e308)) { /* inf or NaN */
    *s = x - x;
    *c = x - x;
    return;
  }
  double r;
  int n = 0;
  if (fabs(x) <= 0.7853981633974483) r = x;
  else n = pm_rem_pio2(x, &r);
  const double sr = pm_ksin(r), cr = pm_kcos(r);
  switch (n) {
    case 0: *s = sr; *c = cr; break;
    case 1: *s = cr; *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
  }
}
PM_FN double pm_sin(double x) {
  double s, c;
  pm_sincos(x, &s, &c);
  return s;
}
PM_FN double pm_cos(double x) {
  double s, c;
  pm_sincos(x, &s, &c);
  return c;
}
PM_FN double pm_tan(double x) {
  double s, c;
  pm_sincos(x, &s, &c);
  return s / c;
}

/* log(x): x = 2^k * m, m in [sqrt(2)/2, sqrt(2)); f = m - 1, s = f / (2 + f), log(m) = f - (f*f/2 - s*(f*f/2 + R)),
 * R = z*(2/3 + z*(2/5 + ...)) with z = s*s (the series of 2*atanh(s) - 2*s, twelve terms: |s| <= 0.1716). */
PM_FN double pm_log(double x) {
  const double ln2_hi = 0.6931467056274414;      /* 0x1.62e4200000000p-1 : k * ln2_hi is exact */
  const double ln2_lo = 4.7493250390316726e-07;  /* 0x1.fdf473de6af28p-22 */
  if (!(x > 0.0)) {
    if (x == 0.0) return -HUGE_VAL;
    return (x - x) / (x - x); /* negative or NaN */
  }
  if (!(x <= 1.7976931348623157e308)) return x; /* +inf */
  int k = 0;
  uint64_t u = pm_bits(x);
  if ((u >> 52) == 0) { /* subnormal: scale by 2^54 */
    x = x * 18014398509481984.0;
    u = pm_bits(x);
    k = -54;
  }
  k += (int)(u >> 52) - 1023;
  u = (u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL; /* m in [1, 2) */
  double m = pm_from_bits(u);
  if (m > 1.4142135623730951) {
    m = 0.5 * m;
    k += 1;
  }
  const double f = m - 1.0;
  const double s = f / (2.0 + f);
  const double z = s * s;
  double R = 0.08;
  R = R * z + 0.08695652173913043;
  R = R * z + 0.09523809523809523;
  R = R * z + 0.10526315789473684;
  R = R * z + 0.11764705882352941;
  R = R * z + 0.13333333333333333;
  R = R * z + 0.15384615384615385;
  R = R * z + 0.18181818181818182;
  R = R * z + 0.2222222222222222;
  R = R * z + 0.2857142857142857;
  R = R * z + 0.4;
  R = R * z + 0.6666666666666666;
  R = R * z;
  const double hfsq = 0.5 * f * f;
  const double dk = (double)k;
  return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

/* hypot(x, y) = sqrt(x*x + y*y), rescaled by a power of two when the squares would overflow or underflow. */
PM_FN double pm_hypot(double x, double y) {
  double a = fabs(x), b = fabs(y);
  if (!(a <= 1.7976931348623157e308) || !(b <= 1.7976931348623157e308)) {
    if (a > 1.7976931348623157e308 || b > 1.7976931348623157e308) return HUGE_VAL; /* an infinite side wins over NaN */
    return a + b;                                                                   /* NaN */
  }
  if (a < b) {
    const double t = a;
    a = b;
    b = t;
  }
  if (a == 0.0) return 0.0;
  if (a > 1e150) {
    a = a * 7.458340731200207e-155; /* 2^-512 */
    b = b * 7.458340731200207e-155;
    return sqrt(a * a + b * b) * 1.3407807929942597e+154; /* 2^512 */
  }
  if (a < 1e-150) {
    a = a * 1.3407807929942597e+154;
    b = b * 1.3407807929942597e+154;
    return sqrt(a * a + b * b) * 7.458340731200207e-155;
  }
  return sqrt(a * a + b * b);
}

#endif /* CILQR_PM_MATH_H_ */
