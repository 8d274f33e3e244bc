// dhj_fastmath.cuh — branch-free FP64 elementary functions for the COS kernels.
//
// Why not libdevice: ncu on the first kernel (profiles/README.md, r01 v1) showed 1.7 non-FP64
// instructions per FP64 instruction — 64-bit polynomial constants rebuilt with UMOV pairs, integer
// range-reduction and slow-path branches (Payne-Hanek, denormal division/sqrt fix-ups) that this path
// never needs — and an 8 300-instruction body that thrashes the instruction cache.  The arguments on
// this path are bounded (|angle| < 2^20, exponents in [-745, 709], ratios of normal numbers), so the
// functions below are straight-line code: FMA polynomials with coefficients in __constant__ memory
// (operands come straight from the constant bank), magic-number rounding, integer exponent tricks.
//
// The characteristic function uses table-driven variants (log_tab, exp_tab, atan2_tab: 2.6 KB of tables that the
// kernels copy into shared memory) with shorter polynomials; the highest polynomial coefficients are rounded to
// a high word so that they fit an FP64 instruction's 32-bit immediate (scripts/gen_poly.py imm).
//
// Accuracy (tests/test_fastmath.py, against mpmath): sincos <= 1.5 ulp for |x| <= 1e5, exp / exp_tab <= 1 ulp,
// log_ratio / log_tab abs error <= 2.3e-16 * max(1, |log|), atan2 / atan2_tab <= 2 ulp, rcp/div/sqrt <= 1 ulp,
// rsqrt <= 1.5 ulp.
// NaN propagates; +-inf and out-of-range arguments give the IEEE limits where the path can produce
// them (exp), otherwise NaN.
//
// Compiles with nvcc (device only) and with g++ (tests/host_emu, tests/fastmath check): on the host the
// MUFU seeds are replaced by exact 1/x, 1/sqrt(x) rounded to 20 bits so the Newton steps are exercised.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define DHJ_FM __device__ __forceinline__
#define DHJ_CONSTANT __constant__
#else
#define DHJ_FM inline
#define DHJ_CONSTANT static const
#endif

namespace dhj {
namespace fm {

// ---- bit access --------------------------------------------------------------------------------
DHJ_FM int lo32(double x) {
#if defined(__CUDA_ARCH__)
  return __double2loint(x);
#else
  uint64_t b; memcpy(&b, &x, 8); return (int)(uint32_t)b;
#endif
}
DHJ_FM int hi32(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  uint64_t b; memcpy(&b, &x, 8); return (int)(uint32_t)(b >> 32);
#endif
}
DHJ_FM double from_hilo(int hi, int lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &b, 8); return x;
#endif
}
// Comparisons on the INTEGER pipe.  Time on this path follows the FP64 instruction count (profiles/README.md r02), and
// a DSETP holds an FP64 issue slot like a DFMA: where a comparison only needs the sign or the order of two
// non-negative numbers, the bit patterns say the same.  sign_bit(-0.0) and sign_bit(-NaN) are true; callers note why
// that is harmless where they use it.
DHJ_FM bool sign_bit(double x) { return hi32(x) < 0; }
DHJ_FM bool gt_nonneg(double a, double b) {            // a > b for a, b >= 0 (a NaN orders above everything)
#if defined(__CUDA_ARCH__)
  return __double_as_longlong(a) > __double_as_longlong(b);
#else
  int64_t ia, ib; memcpy(&ia, &a, 8); memcpy(&ib, &b, 8); return ia > ib;
#endif
}

// x with its sign bit xor-ed with bit 31 of `signmask`
DHJ_FM double xor_sign(double x, int signmask) { return from_hilo(hi32(x) ^ (signmask & (int)0x80000000), lo32(x)); }

// ---- reciprocal, division, square roots ----------------------------------------------------------
// 20-bit seeds: MUFU.RCP64H / MUFU.RSQ64H on the device
DHJ_FM double rcp_seed(double x) {
#if defined(__CUDA_ARCH__)
  double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r;
#else
  return from_hilo(hi32(1.0 / x), 0);
#endif
}
DHJ_FM double rsqrt_seed(double x) {
#if defined(__CUDA_ARCH__)
  double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r;
#else
  return from_hilo(hi32(1.0 / sqrt(x)), 0);
#endif
}

// 1/x for normal x: one cubic step from the seed, r (1 + e + e^2) with e = 1 - x r, takes 20 bits to 60
// (<= 1 ulp; `div` below repairs the last bit with the exact residual)
DHJ_FM double rcp(double x) {
  const double r = rcp_seed(x);
  const double e = fma(-x, r, 1.0);
  return fma(fma(e, e, e), r, r);
}

// a/b for normal operands and quotient (<= 1 ulp, nearly always correctly rounded)
DHJ_FM double div(double a, double b) {
  const double r = rcp(b);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}

// 1/sqrt(x) and sqrt(x) for normal x > 0 (sqrt(0) = 0 handled; <= 1 ulp)
// one third-order step: y (1 + e/2 + 3 e^2/8) with e = 1 - x y^2, 20 bits -> 60
DHJ_FM double rsqrt(double x) {
  const double y = rsqrt_seed(x);
  const double e = fma(-(x * y), y, 1.0);
  return fma(y * e, fma(0.375, e, 0.5), y);
}
DHJ_FM void sqrt_rsqrt(double x, double* s_out, double* y_out) {
  const double y = rsqrt(x);
  double s = x * y;
  s = fma(fma(-s, s, x), 0.5 * y, s);
  *s_out = (x == 0.0) ? 0.0 : s;
  *y_out = y;
}
DHJ_FM double sqrt_(double x) { double s, y; sqrt_rsqrt(x, &s, &y); return s; }

// scalar constants live in constant memory too: as literals every use costs two IMAD.MOV (2 issue cycles each
// on B200) to build the 64-bit immediate; from the constant bank it is one load
struct ScalarConsts {
  double TwoOverPi;
  double Pio2Hi;
  double Pio2Mid;
  double Log2e;
  double Ln2Hi;
  double Ln2Lo;
  double Ln2HiFull;
  double Ln2LoFull;
  double Sqrt2;
  double SqrtHalf;
  double TanPi8;
  double PiO4;
  double PiO4Lo;
  double PiO2;
  double PiD;
  double Log2e64;      // 64 / ln 2
  double Ln2Hi64;      // (ln 2)/64, high 32 bits and the rest
  double Ln2Lo64;
  double AtanMagic;    // 1.5 * 2^46: (t + magic) - magic = t rounded to a multiple of 1/64, low word = 64 t
  double AtC1, AtC2;   // atan t = t + t z (C1 + C2 z + C3 z^2), z = t^2; C3 = -1/7 is an immediate (atan2_tab_impl)
};
DHJ_CONSTANT ScalarConsts kS = {
  6.36619772367581382433e-01,
  1.57079632679489655800e+00,
  6.12323399573676603587e-17,
  1.44269504088896338700e+00,
  6.93147180369123816490e-01,
  1.90821492927058770002e-10,
  6.93147180559945286227e-01,
  2.31904681384629955842e-17,
  1.41421356237309514547e+00,
  7.07106781186547572737e-01,
  4.14213562373095048802e-01,
  7.85398163397448278999e-01,
  3.06161699786838301793e-17,
  1.57079632679489655800e+00,
  3.14159265358979311600e+00,
  1.44269504088896338700e+00 * 64.0,
  6.93147180369123816490e-01 / 64.0,
  1.90821492927058770002e-10 / 64.0,
  1.5 * 70368744177664.0,
  -1.0 / 3.0, 0.2};

// ---- sincos ----------------------------------------------------------------------------------------
// sin r = r + r z S(z), cos r = 1 - z/2 + z^2 C(z), z = r^2 <= (pi/4)^2: minimax fits (scripts/gen_poly.py imm:
// 2.4e-17 / 1.6e-18) whose two highest coefficients are rounded to a HIGH WORD (low 32 bits zero) with the
// rest refitted around them.  An FP64 instruction on sm_100 takes such a constant as a 32-bit immediate;
// any other coefficient costs a constant-bank load (an issue slot) per use.
DHJ_CONSTANT double kSin[4] = {-0.16666666666666664702, 0.0083333333333307815961, -0.00019841269836439504943,
                               2.7557315924814562483e-6};
constexpr double kSin4 = -2.5051093643924104981e-8, kSin5 = 1.5915335715988021548e-10;      // high-word constants
DHJ_CONSTANT double kCos[4] = {0.041666666666666665048, -0.0013888888888887096165, 0.000024801587298387249032,
                               -2.755731711353240655e-7};
constexpr double kCos4 = 2.0876118611568017513e-9, kCos5 = -1.1380882347644671881e-11;      // high-word constants
constexpr double kRoundMagic = 6755399441055744.0;     // 1.5 * 2^52: (x + magic) - magic = rint(x), low word = int

// sin and cos of x, |x| <= ~1e5 (Cody-Waite with FMA; no Payne-Hanek path)
DHJ_FM void sincos_(double x, double* s_out, double* c_out) {
  const double t = fma(x, kS.TwoOverPi, kRoundMagic);
  const int q = lo32(t);
  const double n = t - kRoundMagic;
  double r = fma(-n, kS.Pio2Hi, x);
  r = fma(-n, kS.Pio2Mid, r);               // |n| < 2^17: the next term of pi/2 (1.5e-33 n) is far below 1 ulp
  const double z = r * r;
  double ps = kSin5;
  ps = fma(ps, z, kSin4); ps = fma(ps, z, kSin[3]); ps = fma(ps, z, kSin[2]); ps = fma(ps, z, kSin[1]);
  ps = fma(ps, z, kSin[0]);
  const double s = fma(r * z, ps, r);
  double pc = kCos5;
  pc = fma(pc, z, kCos4); pc = fma(pc, z, kCos[3]); pc = fma(pc, z, kCos[2]); pc = fma(pc, z, kCos[1]);
  pc = fma(pc, z, kCos[0]);
  const double c = fma(z, fma(z, pc, -0.5), 1.0);
  // quadrant: q odd swaps; bit 1 of q negates sin, bit 1 of (q+1) negates cos
  const bool swap = (q & 1) != 0;
  const double ss = swap ? c : s;
  const double cc = swap ? s : c;
  *s_out = xor_sign(ss, q << 30);
  *c_out = xor_sign(cc, (q + 1) << 30);
}
DHJ_FM double cos_(double x) { double s, c; sincos_(x, &s, &c); return c; }

// ---- exp -----------------------------------------------------------------------------------------
// exp(r) = 1 + r + r^2 q(r) on |r| <= ln2/2, q minimax of degree 9 (scripts/gen_poly.py: 1.0e-16)
DHJ_CONSTANT double kExpQ[10] = {0.50000000000000010212, 0.16666666666666674523, 0.041666666666624157875,
                                 0.0083333333333222152106, 0.001388888891719838631, 0.00019841269886566398025,
                                 0.00002480152132015615238, 2.7557242364849317814e-6, 2.7620076799169896683e-7,
                                 2.5110039179932520501e-8};

// core: p * 2^n with p in [0.70, 1.42] and 2^n built in the exponent field.  The scaling is a multiplication, not
// an integer add into p's exponent, so that a NaN p stays NaN whatever bits n carries.  Valid for
// -708 < x < 709.08 (n in [-1022, 1023]); callers guard the rest.
DHJ_FM double exp_core(double x) {
  const double t = fma(x, kS.Log2e, kRoundMagic);
  const int n = lo32(t);
  const double nf = t - kRoundMagic;
  double r = fma(-nf, kS.Ln2Hi, x);
  r = fma(-nf, kS.Ln2Lo, r);
  double q = kExpQ[9];
  q = fma(q, r, kExpQ[8]); q = fma(q, r, kExpQ[7]); q = fma(q, r, kExpQ[6]); q = fma(q, r, kExpQ[5]);
  q = fma(q, r, kExpQ[4]); q = fma(q, r, kExpQ[3]); q = fma(q, r, kExpQ[2]); q = fma(q, r, kExpQ[1]);
  q = fma(q, r, kExpQ[0]);
  const double p = 1.0 + fma(r * r, q, r);
  return p * from_hilo((n + 1023) << 20, 0);
}
// full range
DHJ_FM double exp_(double x) {
  double res = exp_core(x);
  res = (x < -708.0) ? 0.0 : res;          // results below the smallest normal are flushed to zero
  res = (x > 709.08) ? (double)INFINITY : res;    // (e^709.08 = 8.9e307: the last 0.7 of the range overflows early)
  return res;
}
// for arguments known to be <= ~700 (decay factors): only the underflow side is guarded
DHJ_FM double exp_neg(double x) {
  const double res = exp_core(x);
  return (x < -708.0) ? 0.0 : res;
}

// ---- log of a ratio --------------------------------------------------------------------------------
// log(a/b) for normal a, b > 0:  a/b = 2^k * m, m in [sqrt(1/2), sqrt(2));  s = (a - b')/(a + b') with
// b' = b * 2^k;  log m = 2 atanh(s) = 2s + s R(s^2)   (fdlibm's Lg1..Lg7, error 2^-58.45)
DHJ_CONSTANT double kLg[7] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
                              2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
                              1.479819860511658591e-01};

DHJ_FM double log_ratio(double a, double b) {
  // align exponents: b' = b * 2^(ea - eb) has the exponent of a, so a/b' in (1/2, 2)
  int k = ((hi32(a) >> 20) & 0x7ff) - ((hi32(b) >> 20) & 0x7ff);
  double bs = from_hilo(hi32(b) + (k << 20), lo32(b));
  const bool up = a > kS.Sqrt2 * bs, down = a < kS.SqrtHalf * bs;
  bs = up ? 2.0 * bs : (down ? 0.5 * bs : bs);
  k += up ? 1 : (down ? -1 : 0);
  const double s = div(a - bs, a + bs);
  const double z = s * s;
  double R = kLg[6];
  R = fma(R, z, kLg[5]); R = fma(R, z, kLg[4]); R = fma(R, z, kLg[3]); R = fma(R, z, kLg[2]); R = fma(R, z, kLg[1]);
  R = fma(R, z, kLg[0]);
  R = R * z;
  const double kf = (double)k;
  // k ln2 + 2s + s R, low-order parts first
  const double res = fma(kf, kS.Ln2HiFull, fma(s, R, fma(kf, kS.Ln2LoFull, s + s)));
  const double nan_probe = a + b;          // the exponent surgery above would launder a NaN b
  return (nan_probe != nan_probe) ? nan_probe : res;
}
DHJ_FM double log_(double x) { return log_ratio(x, 1.0); }

// ---- table-driven log ------------------------------------------------------------------------------
// log(w) = e ln2 - log(r_i) + log1p(f r_i - 1) for w = 2^e f, f in [1,2), i = top 6 mantissa bits, r_i ~ 1/f:
// |f r_i - 1| <= 2^-7, so a degree-7 Taylor polynomial of log1p is exact to 2e-18.  No division (log_ratio spends
// 8 FP64 instructions on one) and no exponent alignment: 12 FP64 + ~7 integer instructions.  Absolute error
// <= 2e-16 max(1, |log w|).  w must be a positive normal number; NaN / inf give NaN.
struct LogEntry { double r, l; };
// both lookup tables (1.5 KB); read with per-lane indices, so the kernels keep a copy in shared memory
struct Tables {
  LogEntry log[64];        // log_tab: {r_i, -log r_i}
  double exp2[64];         // exp_tab: 2^(j/64)
  double atan64[65];       // atan2_tab: atan(i/64)
};
#if defined(__CUDACC__)
__device__ const Tables kTables = {{
#else
static const Tables kTables = {{
#endif
#include "dhj_logtable.inc"
}, {
#include "dhj_exptable.inc"
}, {
#include "dhj_atantable.inc"
}};

// e_off: returns log(w * 2^e_off) (the offset merges with the exponent bias: free)
// NAN_GUARD = false drops the final `+ (w - w)` (2 FP64 instructions): for callers whose result is NaN anyway when w is.
template <bool NAN_GUARD = true>
DHJ_FM double log_tab(double w, const Tables* __restrict__ tab, int e_off = 0) {
  const int hi = hi32(w);
  const int e = ((hi >> 20) & 0x7ff) - 1023 + e_off;
  const LogEntry t = tab->log[(hi >> 14) & 63];
  const double f = from_hilo((hi & 0x000fffff) | 0x3ff00000, lo32(w));
  const double ep = fma(f, t.r, -1.0);
  // 1/7 and -1/6 rounded to a high word (32-bit immediates): their terms are below 2^-42 and 2^-35 of the result
  double q = 0.14285719394683837891;
  q = fma(q, ep, -0.16666662693023681641); q = fma(q, ep, 0.2); q = fma(q, ep, -0.25); q = fma(q, ep, 1.0 / 3.0);
  q = fma(q, ep, -0.5);
  const double l1p = fma(ep * ep, q, ep);
  const double ef = (double)e;
  // ln 2 - Ln2HiFull rounded to a high word (32-bit immediate): the 21-bit constant misses 1e-23 |e|
  const double res = fma(ef, kS.Ln2HiFull, t.l) + fma(ef, 2.31904650766222601363e-17, l1p);
  return NAN_GUARD ? res + (w - w) : res;   // NaN or inf in -> NaN out (the bit surgery above would launder them)
}

// ---- table-driven exp ------------------------------------------------------------------------------
// exp(x) = 2^m * 2^(j/64) * exp(r), n = rint(64 x / ln 2) = 64 m + j, |r| <= ln2/128: the polynomial shrinks from
// degree 11 to 5 (r + r^2 q(r), q minimax of degree 3 with a high-word leading coefficient: 5e-18), 11 FP64
// instructions instead of 17.  <= 1 ulp.
// Same range rules as exp_core / exp_ / exp_neg.
DHJ_CONSTANT double kExpT[3] = {0.49999999999985997477, 0.1666666666666058649, 0.041666707400273529831};
constexpr double kExpT3 = 0.0083333402872085571289;                                           // high-word constant

DHJ_FM double exp_tab_core(double x, const Tables* __restrict__ tab) {
  const double t = fma(x, kS.Log2e64, kRoundMagic);
  const int n = lo32(t);
  const double nf = t - kRoundMagic;
  double r = fma(-nf, kS.Ln2Hi64, x);
  r = fma(-nf, kS.Ln2Lo64, r);
  const double s = tab->exp2[n & 63];
  double q = kExpT3;
  q = fma(q, r, kExpT[2]); q = fma(q, r, kExpT[1]); q = fma(q, r, kExpT[0]);
  const double p = fma(r * r, q, r);
  return fma(s, p, s) * from_hilo(((n >> 6) + 1023) << 20, 0);
}
DHJ_FM double exp_tab(double x, const Tables* __restrict__ tab) {
  double res = exp_tab_core(x, tab);
  res = (x < -708.0) ? 0.0 : res;
  res = (x > 709.08) ? (double)INFINITY : res;
  return res;
}
DHJ_FM double exp_tab_neg(double x, const Tables* __restrict__ tab) {
  const double res = exp_tab_core(x, tab);
  return (x < -708.0) ? 0.0 : res;
}
// the same with the underflow test on the integer pipe: the high word of a double below -708 is, as an unsigned
// number, above that of -708.0 (0xC0862000); -inf flushes to 0 as it must, and so does a NEGATIVE NaN — for callers
// whose result is NaN anyway when x is
DHJ_FM double exp_tab_neg_ix(double x, const Tables* __restrict__ tab) {
  const double res = exp_tab_core(x, tab);
  return ((unsigned)hi32(x) > 0xC0862000u) ? 0.0 : res;
}

// ---- atan2 -----------------------------------------------------------------------------------------
// atan(t) = t - t z P(z), z = t^2, |t| <= tan(pi/8)  (fdlibm aT[0..10], designed for |t| < 7/16)
DHJ_CONSTANT double kAt[11] = {3.33333333333329318027e-01, -1.99999999998764832476e-01, 1.42857142725034663711e-01,
                               -1.11111104054623557880e-01, 9.09088713343650656196e-02, -7.69187620504482999495e-02,
                               6.66107313738753120669e-02, -5.83357013379057348645e-02, 4.97687799461593236017e-02,
                               -3.65315727442169155270e-02, 1.62858201153657823623e-02};

// ZERO_OK = false drops the atan2(+-0, +-0) special case (the hot path's argument D conj(d) is never zero; a zero
// would give NaN): it costs a data-dependent branch pair per call otherwise.
template <bool ZERO_OK>
DHJ_FM double atan2_impl(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  const bool steep = ay > ax;                       // NaN-safe: a NaN operand lands in mn or mx and propagates
  const double mx = steep ? ay : ax, mn = steep ? ax : ay;
  // octant reduction without a second division: atan(mn/mx) = pi/4 + atan((mn-mx)/(mn+mx)) when mn/mx > tan(pi/8)
  const bool hi = mn > kS.TanPi8 * mx;
  const double num = hi ? mn - mx : mn;
  const double den = hi ? mn + mx : mx;
  const double t = div(num, den);
  const double z = t * t;
  double p = kAt[10];
  p = fma(p, z, kAt[9]); p = fma(p, z, kAt[8]); p = fma(p, z, kAt[7]); p = fma(p, z, kAt[6]); p = fma(p, z, kAt[5]);
  p = fma(p, z, kAt[4]); p = fma(p, z, kAt[3]); p = fma(p, z, kAt[2]); p = fma(p, z, kAt[1]); p = fma(p, z, kAt[0]);
  double r = fma(-(t * z), p, t);
  r = hi ? (r + kS.PiO4Lo) + kS.PiO4 : r;     // atan(mn/mx) in [0, pi/4]
  r = steep ? kS.PiO2 - r : r;              // first quadrant angle
  r = (x < 0.0) ? kS.PiD - r : r;
  if (ZERO_OK) r = (mx == 0.0) ? ((x < 0.0 || (x == 0.0 && signbit(x))) ? kS.PiD : 0.0) : r;   // atan2(+-0, +-0)
  return copysign(r, y);
}
DHJ_FM double atan2_(double y, double x) { return atan2_impl<true>(y, x); }
DHJ_FM double atan2_nz(double y, double x) { return atan2_impl<false>(y, x); }

// table-driven variant: with c = mn/mx rounded to a multiple of 1/64 (from the 20-bit reciprocal seed),
// atan(mn/mx) = atan(c) + atan((mn - c mx)/(mx + c mn)) and the second argument is below 1/127: three terms of
// the series instead of the degree-21 polynomial, and no octant step.  21 FP64 instructions instead of 30;
// <= 2 ulp.  Same special cases as atan2_impl (a NaN operand gives NaN: the table index is clamped).
template <bool ZERO_OK>
DHJ_FM double atan2_tab_impl(double y, double x, const Tables* __restrict__ tab) {
  const double ax = fabs(x), ay = fabs(y);
  const bool steep = gt_nonneg(ay, ax);             // integer compare; a NaN operand lands in mx or mn either way
  const double mx = steep ? ay : ax, mn = steep ? ax : ay;
  const double tt = fma(mn, rcp_seed(mx), kS.AtanMagic);
  unsigned i = (unsigned)lo32(tt);
  i = i < 64u ? i : 64u;
  const double c = tt - kS.AtanMagic;
  // |t| < 1/127: the quotient's last-bit repair (fm::div) would change the result by < 2^-60; one product suffices
  const double t = fma(-c, mx, mn) * rcp(fma(c, mn, mx));
  const double z = t * t;
  const double p = fma(z, fma(z, -0.14285719394683837891, kS.AtC2), kS.AtC1);   // -1/7 as a high word: immediate
  double r = tab->atan64[i] + fma(t * z, p, t);     // atan(mn/mx) in [0, pi/4]
  r = steep ? kS.PiO2 - r : r;
  // sign bit instead of x < 0: differs for x = -0.0 only, where r = pi/2 exactly and pi - pi/2 = pi/2 exactly
  r = sign_bit(x) ? kS.PiD - r : r;
  if (ZERO_OK) r = (mx == 0.0) ? ((x < 0.0 || (x == 0.0 && signbit(x))) ? kS.PiD : 0.0) : r;
  return copysign(r, y);
}
DHJ_FM double atan2_tab(double y, double x, const Tables* __restrict__ tab) { return atan2_tab_impl<true>(y, x, tab); }
DHJ_FM double atan2_tab_nz(double y, double x, const Tables* __restrict__ tab) { return atan2_tab_impl<false>(y, x, tab); }

}  // namespace fm
}  // namespace dhj
