/* oracle/rngmath.h -- the two transcendental maps of the RNG specification (DESIGN.md "RNG"): uniforms -> exponential
 * variates (rng_log) and -> Box-Muller angles (rng_sincos2pi).  TEST INFRASTRUCTURE (see oracle.h).
 *
 * These are part of the SPECIFICATION of the Philox streams, not a restatement of reference arithmetic (the reference draws
 * from R's Mersenne-Twister stream through norm_rand / exp_rand, which no GPU code can reproduce).  Each function is a fixed
 * sequence of IEEE-754 double operations (+, *, fma, rint, integer bit operations): the CUDA library executes the same
 * sequence with __dadd_rn / __dmul_rn / __fma_rn (csrc/common.cuh), so a uniform maps to the SAME double on both sides.
 * Accuracy (tests/test_oracle_primitives.py): |rng_log(u) - log(u)| <= 4e-16 |log(u)|, sin/cos within 3e-16 absolute.
 * Model arithmetic that restates the reference (densities, transitions) keeps libm log / exp / cos. */
#ifndef CPIMH_ORACLE_RNGMATH_H
#define CPIMH_ORACLE_RNGMATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

/* log(u) for a positive normal double u < 2: u = 2^e m, m in (sqrt(1/2), sqrt(2)]; s = (m-1)/(m+1) by a Newton reciprocal;
 * log m = 2 atanh(s) = 2s (1 + z/3 + ... + z^10/21), z = s^2 <= 0.0295 */
static inline double orc_rng_log(double u) {
  uint64_t b;
  memcpy(&b, &u, 8);
  int e = (int)(b >> 52) - 1023;
  b = (b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
  double m;
  memcpy(&m, &b, 8);
  if (m > 0x1.6a09e667f3bcdp+0) { m = m * 0.5; e = e + 1; }
  const double f = m - 1.0, d = m + 1.0;
  const double h = d * 0.5;                       /* in [0.853, 1.208] */
  const double t = 1.0 - h;
  double r = fma(t, 1.0 + t, 1.0);                /* 1/h to 1e-2, then three Newton steps */
  double eps = fma(-h, r, 1.0);
  r = fma(r, eps, r);
  eps = fma(-h, r, 1.0);
  r = fma(r, eps, r);
  eps = fma(-h, r, 1.0);
  r = fma(r, eps, r);
  const double rd = r * 0.5;                      /* 1/d */
  double s = f * rd;
  const double rem = fma(-s, d, f);
  s = fma(rem, rd, s);                            /* s = f/d */
  const double z = s * s;
  double p = 0x1.8618618618618p-5;                /* 1/21 */
  p = fma(p, z, 0x1.af286bca1af28p-5);            /* 1/19 */
  p = fma(p, z, 0x1.e1e1e1e1e1e1ep-5);            /* 1/17 */
  p = fma(p, z, 0x1.1111111111111p-4);            /* 1/15 */
  p = fma(p, z, 0x1.3b13b13b13b14p-4);            /* 1/13 */
  p = fma(p, z, 0x1.745d1745d1746p-4);            /* 1/11 */
  p = fma(p, z, 0x1.c71c71c71c71cp-4);            /* 1/9  */
  p = fma(p, z, 0x1.2492492492492p-3);            /* 1/7  */
  p = fma(p, z, 0x1.999999999999ap-3);            /* 1/5  */
  p = fma(p, z, 0x1.5555555555555p-2);            /* 1/3  */
  const double s2 = s + s;
  const double lm = fma(s2 * z, p, s2);
  return fma((double)e, 0x1.62e42fefa39efp-1, lm);
}

/* sin(2 pi a), cos(2 pi a) for a in [0, 1]: a = q/4 + r, |r| <= 1/8, Taylor series of sin/cos(2 pi r) in r^2, quadrant fix-up */
static inline void orc_rng_sincos2pi(double a, double* sn, double* cs) {
  const double q = rint(4.0 * a);
  const double r = fma(-0.25, q, a);
  const double z = r * r;
  double ps = 0x1.aaec32af93359p-4;
  ps = fma(ps, z, -0x1.6fadb9f155744p-1);
  ps = fma(ps, z, 0x1.e8f434d018d63p+1);
  ps = fma(ps, z, -0x1.e3074fde8871fp+3);
  ps = fma(ps, z, 0x1.50783487ee782p+5);
  ps = fma(ps, z, -0x1.32d2cce62bd86p+6);
  ps = fma(ps, z, 0x1.466bc6775aae2p+6);
  ps = fma(ps, z, -0x1.4abbce625be53p+5);
  ps = fma(ps, z, 0x1.921fb54442d18p+2);
  const double sr = r * ps;
  double pc = -0x1.2a0c591af8314p-5;
  pc = fma(pc, z, 0x1.20c62c2f2d7f5p-2);
  pc = fma(pc, z, -0x1.b6e24f44b128fp+0);
  pc = fma(pc, z, 0x1.f9d38a3763cc3p+2);
  pc = fma(pc, z, -0x1.a6d1f2a204a8cp+4);
  pc = fma(pc, z, 0x1.e1f506891babbp+5);
  pc = fma(pc, z, -0x1.55d3c7e3cbffap+6);
  pc = fma(pc, z, 0x1.03c1f081b5ac4p+6);
  pc = fma(pc, z, -0x1.3bd3cc9be45dep+4);
  const double cr = fma(z, pc, 1.0);
  const int qi = (int)q & 3;
  *sn = (qi & 1) ? cr : sr;
  *cs = (qi & 1) ? sr : cr;
  if (qi == 1 || qi == 2) *cs = -*cs;
  if (qi >= 2) *sn = -*sn;
}
#endif
