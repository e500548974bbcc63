// lgar_c_b200/csrc/lgar_pow.cuh
//
// x^y that returns THE SAME double as the C library the reference links against.
//
// The reference evaluates every van Genuchten relation through libm pow() (reference src/soil_funcs.cxx:147-150,155-159,
// 164-167,172-179; src/aet.cxx:60; src/conceptual_reservoir.cxx:19-35) and its front logic is threshold-triggered: a last-bit
// difference in one pow can end a psi search one trip apart, move 1e-10 cm between AET and runoff, or merge a front one step
// earlier (with exp(y log x) 59 of 10 000 config-4 columns left the 1e-9 bar and one config-5 column split
// its front count).  A pow that is merely accurate does not close that; one that rounds the way the host's does, does.
//
// The host function is the table-driven pow of glibc >= 2.28 (Szabolcs Nagy's algorithm, published in ARM optimized-routines
// pow.c): log(x) is carried as hi + lo with ~2^-68 relative error from a 128-entry {1/c, log c} table and a degree-7 polynomial,
// y*log(x) is formed as ehi + elo with one fma, and exp is 2^(k/128) * (1 + tail + P(r)) from a 128-entry table.  What is restated
// here is that algorithm as the x86-64 FMA build executes it (the variant the ifunc resolver picks on every x86 host with
// FMA + AVX2, i.e. every GPU host): which products are fused into an fma was read off the instruction stream of that build, and is
// written out below as explicit fma() calls -- the file is compiled with -fmad=false, so nothing else contracts.  Special cases
// (zero, negative, infinite, NaN, subnormal bases; huge and tiny exponents; overflow, underflow into the subnormal range) follow
// the same published source.  errno and floating-point exception flags are not reproduced (the reference reads neither).
//
// Checked bit for bit against the host pow (tests/test_pow_exact.py: 2e8 random (x, y) of the shapes the path produces, the whole
// special-case lattice, and random bit patterns), on the CPU through LGAR_HOSTSIM and on the GPU through lgarb_debug_pow.
#ifndef LGAR_POW_CUH
#define LGAR_POW_CUH

#include <stdint.h>

#include "lgar_pow_tables.h"

#ifdef LGAR_HOSTSIM
#include <cmath>
#include <cstring>
#define LGP_FN static inline
#define LGP_TABLE static const
static inline uint64_t lgp_bits(double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; }
static inline double lgp_double(uint64_t u) { double x; std::memcpy(&x, &u, 8); return x; }
static inline double lgp_fma(double a, double b, double c) { return std::fma(a, b, c); }
#define LGP_LDG(p) (*(p))
#else
#define LGP_FN __device__ __forceinline__
#define LGP_TABLE static __device__ const
LGP_FN uint64_t lgp_bits(double x) { return (uint64_t)__double_as_longlong(x); }
LGP_FN double lgp_double(uint64_t u) { return __longlong_as_double((long long)u); }
LGP_FN double lgp_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
#define LGP_LDG(p) __ldg(p)
#endif

// 3 KB + 2 KB, read through the read-only path: every lane of a warp indexes its own entry, so __constant__ (one address per
// warp instruction) would serialise; the L1 holds both tables.
LGP_TABLE double lgp_log_tab[128 * 3] = {
    LGP_LOG_TABLE
};
LGP_TABLE unsigned long long lgp_exp_tab[256] = {LGP_EXP_TABLE};

#define LGP_SIGN_BIAS (0x800u << 7)

// 0 if y is not an integer, 1 if odd, 2 if even
LGP_FN int lgp_checkint(uint64_t iy) {
  const int e = (int)(iy >> 52) & 0x7ff;
  if (e < 0x3ff) return 0;
  if (e > 0x3ff + 52) return 2;
  if (iy & ((1ULL << (0x3ff + 52 - e)) - 1)) return 0;
  if (iy & (1ULL << (0x3ff + 52 - e))) return 1;
  return 2;
}
LGP_FN bool lgp_zeroinfnan(uint64_t i) { return 2 * i - 1 >= 2 * 0x7ff0000000000000ULL - 1; }

// the result scaled into / out of range when 2^(k/N) itself is not a valid double
LGP_FN double lgp_exp_special(double tmp, uint64_t sbits, uint64_t ki) {
  if ((ki & 0x80000000u) == 0) {
    sbits -= 1009ULL << 52;
    const double scale = lgp_double(sbits);
    return 0x1p1009 * lgp_fma(scale, tmp, scale);
  }
  sbits += 1022ULL << 52;
  const double scale = lgp_double(sbits);
  const double st = scale * tmp;
  double y = scale + st;
  if (fabs(y) < 1.0) {
    double one = 1.0;
    if (y < 0.0) one = -1.0;
    double lo = scale - y + st;
    const double hi = one + y;
    lo = one - hi + y + lo;
    y = (hi + lo) - one;
    if (y == 0.0) y = lgp_double(sbits & 0x8000000000000000ULL);
  }
  return 0x1p-1022 * y;
}

// exp(x + xtail) * (-1)^(sign_bias != 0)
LGP_FN double lgp_exp_inline(double x, double xtail, uint32_t sign_bias) {
  uint32_t abstop = (uint32_t)(lgp_bits(x) >> 52) & 0x7ff;
  if (abstop - 0x3c9u >= 0x3fu) {
    if (abstop - 0x3c9u >= 0x80000000u) {
      const double one = 1.0 + x;
      return sign_bias ? -one : one;
    }
    if (abstop >= 0x409u) {
      if (lgp_bits(x) >> 63) return sign_bias ? -0.0 : 0.0;                  // underflow
      return sign_bias ? -lgp_double(0x7ff0000000000000ULL) : lgp_double(0x7ff0000000000000ULL);  // overflow
    }
    abstop = 0;
  }
  double kd = lgp_fma(x, LGP_INVLN2N, LGP_SHIFT);
  const uint64_t ki = lgp_bits(kd);
  kd -= LGP_SHIFT;
  double r = lgp_fma(kd, LGP_NEGLN2LON, lgp_fma(kd, LGP_NEGLN2HIN, x));
  r += xtail;
  const uint32_t idx = 2u * (uint32_t)(ki & 127u);
  const uint64_t top = (ki + sign_bias) << 45;
  const double tail = lgp_double(LGP_LDG(&lgp_exp_tab[idx]));
  const uint64_t sbits = LGP_LDG(&lgp_exp_tab[idx + 1]) + top;
  const double r2 = r * r;
  const double p23 = lgp_fma(r, LGP_C3, LGP_C2);
  const double p45 = lgp_fma(r, LGP_C5, LGP_C4);
  const double tmp = lgp_fma(p45, r2 * r2, lgp_fma(p23, r2, tail + r));
  if (abstop == 0) return lgp_exp_special(tmp, sbits, ki);
  const double scale = lgp_double(sbits);
  return lgp_fma(scale, tmp, scale);
}

// log(x) = hi + *tail for the bit pattern ix of a positive normal(ised) x
LGP_FN double lgp_log_inline(uint64_t ix, double* tail) {
  const uint64_t tmp = ix - 0x3fe6955500000000ULL;
  const int i = (int)((tmp >> 45) & 127u);
  const int k = (int)((int64_t)tmp >> 52);
  const uint64_t iz = ix - (tmp & (0xfffULL << 52));
  const double z = lgp_double(iz);
  const double kd = (double)k;
  const double invc = LGP_LDG(&lgp_log_tab[3 * i]);
  const double logc = LGP_LDG(&lgp_log_tab[3 * i + 1]);
  const double logctail = LGP_LDG(&lgp_log_tab[3 * i + 2]);
  const double r = lgp_fma(z, invc, -1.0);
  const double t1 = lgp_fma(kd, LGP_LN2HI, logc);
  const double t2 = t1 + r;
  const double lo1 = lgp_fma(kd, LGP_LN2LO, logctail);
  const double lo2 = t1 - t2 + r;
  const double ar = LGP_A0 * r;
  const double ar2 = r * ar;
  const double ar3 = r * ar2;
  const double hi = t2 + ar2;
  const double lo3 = lgp_fma(ar, r, -ar2);
  const double lo4 = t2 - hi + ar2;
  const double q12 = lgp_fma(r, LGP_A2, LGP_A1);
  const double q34 = lgp_fma(r, LGP_A4, LGP_A3);
  const double q56 = lgp_fma(r, LGP_A6, LGP_A5);
  const double q = lgp_fma(ar2, lgp_fma(q56, ar2, q34), q12);
  const double lo = lgp_fma(ar3, q, lo1 + lo2 + lo3 + lo4);
  const double y = hi + lo;
  *tail = hi - y + lo;
  return y;
}

// The common case only -- positive normal x, 2^-65 <= |y| < 2^63, 2^-54 <= |y log x| < 512 -- for the callers that sit in a hot
// loop and want it inlined (the psi search): false means "take the full function".  Same arithmetic as lgp_pow on that domain.
LGP_FN bool lgp_try_pow(double x, double y, double* out) {
  const uint64_t ix = lgp_bits(x);
  const uint32_t topx = (uint32_t)(ix >> 52);
  const uint32_t topy = (uint32_t)(lgp_bits(y) >> 52);
  if (topx - 0x001u >= 0x7ffu - 0x001u || (topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) return false;
  double lo;
  const double hi = lgp_log_inline(ix, &lo);
  const double ehi = y * hi;
  const double elo = lgp_fma(y, lo, lgp_fma(y, hi, -ehi));
  const uint32_t abstop = (uint32_t)(lgp_bits(ehi) >> 52) & 0x7ff;
  if (abstop - 0x3c9u >= 0x3fu) return false;
  double kd = lgp_fma(ehi, LGP_INVLN2N, LGP_SHIFT);
  const uint64_t ki = lgp_bits(kd);
  kd -= LGP_SHIFT;
  double r = lgp_fma(kd, LGP_NEGLN2LON, lgp_fma(kd, LGP_NEGLN2HIN, ehi));
  r += elo;
  const uint32_t idx = 2u * (uint32_t)(ki & 127u);
  const double tail = lgp_double(LGP_LDG(&lgp_exp_tab[idx]));
  const uint64_t sbits = LGP_LDG(&lgp_exp_tab[idx + 1]) + (ki << 45);
  const double r2 = r * r;
  const double tmp = lgp_fma(lgp_fma(r, LGP_C5, LGP_C4), r2 * r2, lgp_fma(lgp_fma(r, LGP_C3, LGP_C2), r2, tail + r));
  const double scale = lgp_double(sbits);
  *out = lgp_fma(scale, tmp, scale);
  return true;
}

LGP_FN double lgp_pow(double x, double y) {
  uint32_t sign_bias = 0;
  uint64_t ix = lgp_bits(x);
  const uint64_t iy = lgp_bits(y);
  uint32_t topx = (uint32_t)(ix >> 52);
  const uint32_t topy = (uint32_t)(iy >> 52);
  if (topx - 0x001u >= 0x7ffu - 0x001u || (topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
    if (lgp_zeroinfnan(iy)) {
      if (2 * iy == 0) return 1.0;            // (signalling NaNs are not distinguished)
      if (ix == 0x3ff0000000000000ULL) return 1.0;
      if (2 * ix > 2 * 0x7ff0000000000000ULL || 2 * iy > 2 * 0x7ff0000000000000ULL) return x + y;
      if (2 * ix == 2 * 0x3ff0000000000000ULL) return 1.0;
      if ((2 * ix < 2 * 0x3ff0000000000000ULL) == !(iy >> 63)) return 0.0;
      return y * y;
    }
    if (lgp_zeroinfnan(ix)) {
      double x2 = x * x;
      if ((ix >> 63) && lgp_checkint(iy) == 1) x2 = -x2;
      return (iy >> 63) ? 1.0 / x2 : x2;
    }
    if (ix >> 63) {
      const int yint = lgp_checkint(iy);
      if (yint == 0) return lgp_double(0xfff8000000000000ULL);  // (x - x) / (x - x) on x86-64: the default NaN, sign set
      if (yint == 1) sign_bias = LGP_SIGN_BIAS;
      ix &= 0x7fffffffffffffffULL;
      topx &= 0x7ff;
    }
    if ((topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
      if (ix == 0x3ff0000000000000ULL) return 1.0;
      if ((topy & 0x7ff) < 0x3beu) return ix > 0x3ff0000000000000ULL ? 1.0 + y : 1.0 - y;
      return (ix > 0x3ff0000000000000ULL) == (topy < 0x800u) ? lgp_double(0x7ff0000000000000ULL) : 0.0;
    }
    if (topx == 0) {
      ix = lgp_bits(x * 0x1p52);
      ix &= 0x7fffffffffffffffULL;
      ix -= 52ULL << 52;
    }
  }
  double lo;
  const double hi = lgp_log_inline(ix, &lo);
  const double ehi = y * hi;
  const double elo = lgp_fma(y, lo, lgp_fma(y, hi, -ehi));
  return lgp_exp_inline(ehi, elo, sign_bias);
}

#endif
