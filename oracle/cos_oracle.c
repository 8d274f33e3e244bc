/* cos_oracle.c — C99 restatement of the reference's COS pricer and calibration loss.  TEST INFRASTRUCTURE ONLY.
 *
 * Same role and same rules as oracle/cos_oracle.py (only tests/, smoke() and bench.py's CPU-baseline legs may
 * load it; the product never does).  It exists because the NumPy oracle prices ~7 000 options/s: this one,
 * threaded with OpenMP, checks hundreds of thousands of GPU prices in a second and gives bench.py a
 * "best-effort CPU" line next to the reference-style scalar port.
 *
 * It follows the reference operation for operation (/root/reference/src/models/double_heston.py:48-192,
 * /root/reference/src/calibration/lbfgs_calibrator.py:62-177) and reproduces NumPy's complex arithmetic:
 *   - complex * complex : (ac - bd, ad + bc), no FMA (build with -ffp-contract=off);
 *   - complex / complex : Smith's algorithm as in numpy/_core/src/umath/loops (|br| >= |bi| branch etc.);
 *   - z**2 = z*z; sqrt/exp/log of complex = glibc csqrt/cexp/clog (what npymath calls on Linux);
 *   - np.sum = NumPy's pairwise summation (8 accumulators up to 128 elements, recursive halves above).
 * Parity status: PINNED by tests/test_oracle_golden.py::test_c_oracle_* against the fixtures generated from
 * the live reference: 45 % of the 2 250-price grid fixture bit-identical, max relative difference 8e-14, median
 * below 1e-15.  It cannot be bit-identical everywhere: NumPy evaluates REAL exp / sin / cos / log with its own
 * SIMD kernels while this file uses glibc's (single-ulp differences, amplified by the conditioning of the sum).
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef double complex cplx;

static inline cplx c_mul(cplx a, cplx b) {
  const double ar = creal(a), ai = cimag(a), br = creal(b), bi = cimag(b);
  return CMPLX(ar * br - ai * bi, ar * bi + ai * br);
}
static inline cplx c_scale(double s, cplx a) { return c_mul(CMPLX(s, 0.0), a); }   /* float * complex in NumPy */

static inline cplx c_div(cplx a, cplx b) {                 /* Smith, as NumPy */
  const double ar = creal(a), ai = cimag(a), br = creal(b), bi = cimag(b);
  const double abr = fabs(br), abi = fabs(bi);
  if (abr >= abi) {
    if (abr == 0.0 && abi == 0.0) return CMPLX(ar / abr, ai / abi);
    const double rat = bi / br, scl = 1.0 / (br + bi * rat);
    return CMPLX((ar + ai * rat) * scl, (ai - ar * rat) * scl);
  } else {
    const double rat = br / bi, scl = 1.0 / (bi + br * rat);
    return CMPLX((ar * rat + ai) * scl, (ai * rat - ar) * scl);
  }
}

static double pairwise_sum(const double* a, int n) {       /* numpy pairwise_sum_DOUBLE, unit stride */
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res += a[i];
    return res;
  } else if (n <= 128) {
    double r[8];
    int i;
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  } else {
    int n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
  }
}

/* p = (v01, kappa1, theta1, sigma1, rho1, v02, kappa2, theta2, sigma2, rho2, lambda_j, mu_j, sigma_j) */

static void heston_factor(double u, double tau, double v0, double kappa, double theta, double sigma, double rho,
                          cplx* A, cplx* B) {              /* double_heston.py:64-71, 85-87 */
  const double rs = rho * sigma;                           /* kappa - rho*sigma*1j*phi */
  const cplx beta = CMPLX(kappa, 0.0) - CMPLX(rs * 0.0 * u, (rs * 1.0) * u);
  const double s2 = sigma * sigma;                         /* sigma**2 */
  const cplx t2 = c_mul(CMPLX(s2 * u, 0.0), CMPLX(u, 1.0)); /* sigma**2 * phi * (phi + i) */
  const cplx d = csqrt(c_mul(beta, beta) + t2);
  const cplx g = c_div(beta - d, beta + d);
  const cplx E = cexp(c_scale(tau, -d) );                  /* exp(-d * tau) */
  *B = c_mul(c_div(beta - d, CMPLX(s2, 0.0)), c_div(1.0 - E, 1.0 - c_mul(g, E)));
  const double c = kappa * theta / s2;
  const cplx lg = clog(c_div(1.0 - c_mul(g, E), 1.0 - g));
  *A = c_scale(c, c_scale(tau, beta - d) - c_scale(2.0, lg));
}

static cplx cf(double u, double tau, const double* p, double r, double q) {   /* double_heston.py:48-97 */
  cplx A1, B1, A2, B2;
  heston_factor(u, tau, p[0], p[1], p[2], p[3], p[4], &A1, &B1);
  heston_factor(u, tau, p[5], p[6], p[7], p[8], p[9], &A2, &B2);
  const double lam = p[10], mu = p[11], sj = p[12];
  const double comp = exp(mu + 0.5 * (sj * sj)) - 1.0;
  const double drift = r - q - lam * comp;
  cplx A = CMPLX(0.0, (drift * u) * tau);                  /* drift * 1j * phi * tau */
  A = A + A1;
  A = A + A2;
  const cplx je = cexp(CMPLX(-(0.5 * (sj * sj)) * (u * u), u * mu));   /* exp(i phi mu - 0.5 sj^2 phi^2) */
  const cplx cf_jump = cexp(c_scale(lam * tau, je - 1.0));
  const cplx cf_heston = cexp((A + c_scale(p[0], B1)) + c_scale(p[5], B2));
  return c_mul(cf_heston, cf_jump);
}

static void cumulants(double tau, double r, double v0, double lm, double vbar, double vv, double rho, double* c1,
                      double* c2) {                        /* double_heston.py:101-119 */
  const double e = exp(-lm * tau);
  *c1 = r * tau + (1 - e) * (vbar - v0) / (2 * lm) - vbar * tau / 2;
  *c2 = 1 / (8 * pow(lm, 3)) *
        (vv * tau * lm * e * (v0 - vbar) * (8 * lm * rho - 4 * vv) + lm * rho * vv * (1 - e) * (16 * vbar - 8 * v0) +
         2 * vbar * lm * tau * (-4 * lm * rho * vv + pow(vv, 2) + 4 * pow(lm, 2)) +
         pow(vv, 2) * ((vbar - 2 * v0) * exp(-2 * lm * tau) + vbar * (6 * e - 7) + 2 * v0) +
         8 * pow(lm, 2) * (v0 - vbar) * (1 - e));
}

static void truncation_range(const double* p, double S0, double K, double T, double r, double L, double* a,
                             double* b) {                  /* double_heston.py:100-139 */
  double c1a, c2a, c1b, c2b;
  cumulants(T, r, p[0], p[1], p[2], p[3], p[4], &c1a, &c2a);
  cumulants(T, r, p[5], p[6], p[7], p[8], p[9], &c1b, &c2b);
  const double c1 = c1a + c1b + p[10] * T * p[11];
  const double c2 = c2a + c2b + p[10] * T * (p[12] * p[12] + p[11] * p[11]);
  double lo = c1 - L * sqrt(fabs(c2)), hi = c1 + L * sqrt(fabs(c2));
  const double x = log(K / S0);
  if (x - 0.1 < lo) lo = x - 0.1;                          /* Python min / max semantics */
  if (x + 0.1 > hi) hi = x + 0.1;
  *a = lo; *b = hi;
}

static double price_one(const double* p, double S0, double K, double T, double r, double q, int is_call, int N,
                        double L, double* ab) {            /* double_heston.py:160-192 */
  double a, b;
  const double x = log(K / S0);
  truncation_range(p, S0, K, T, r, L, &a, &b);
  if (ab) { ab[0] = a; ab[1] = b; }
  double* terms = (double*)malloc((size_t)N * sizeof(double));
  const double c = is_call ? x : a, d = is_call ? b : x;
  for (int k = 0; k < N; ++k) {
    const double u = (k * M_PI) / (b - a);
    double chi, psi;
    if (k == 0) { chi = exp(d) - exp(c); psi = d - c; }
    else {
      chi = (1.0 / (1 + u * u)) * (cos(u * (d - a)) * exp(d) - cos(u * (c - a)) * exp(c) +
                                   u * sin(u * (d - a)) * exp(d) - u * sin(u * (c - a)) * exp(c));
      psi = (1.0 / u) * (sin(u * (d - a)) - sin(u * (c - a)));
    }
    const double V = is_call ? (2.0 / (b - a)) * (S0 * chi - K * psi) : (2.0 / (b - a)) * (K * psi - S0 * chi);
    const cplx rot = cexp(CMPLX(-0.0 * (u * a), -1.0 * (u * a)));      /* exp(-1j * u * a) */
    terms[k] = creal(c_mul(cf(u, T, p, r, q), rot)) * V;
  }
  terms[0] *= 0.5;
  const double price = exp(-r * T) * pairwise_sum(terms, N);
  free(terms);
  return price;
}

/* ---- exported ------------------------------------------------------------------------------------------- */
void oracle_price_list(const double* params, int64_t P, const double* S0, int64_t s0_stride, const double* strike,
                       int64_t strike_stride, const double* maturity, const int32_t* is_call, int32_t M, double r,
                       double q, int32_t N, double L, double* out, double* ab) {
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t i = 0; i < P * M; ++i) {
    const int64_t p = i / M;
    const int o = (int)(i - p * M);
    out[i] = price_one(params + 13 * p, S0[p * s0_stride], strike[p * strike_stride + o], maturity[o], r, q,
                       is_call[o], N, L, ab ? ab + 2 * i : 0);
  }
}

void oracle_loss_batch(const double* x, int64_t B, double S0, double r, const double* strike, const double* maturity,
                       const int32_t* is_call, const double* market, int32_t M, int32_t N, double* out) {
#pragma omp parallel for schedule(dynamic, 2)
  for (int64_t b = 0; b < B; ++b) {                        /* lbfgs_calibrator.py:62-87, 111-177 */
    double p[13];
    for (int j = 0; j < 13; ++j) p[j] = exp(x[13 * b + j]);
    p[4] = tanh(x[13 * b + 4]); p[9] = tanh(x[13 * b + 9]); p[11] = x[13 * b + 11];
    int bad = 0;
    double* rel2 = (double*)malloc((size_t)M * sizeof(double));
    for (int o = 0; o < M; ++o) {
      const double price = price_one(p, S0, strike[o], maturity[o], r, 0.0, is_call[o], N, 10.0, 0);
      if (isnan(price) || isinf(price) || price <= 0) bad = 1;
      const double rel = (price - market[o]) / market[o];
      rel2[o] = rel * rel;
    }
    const double e1 = p[3] * p[3] - 2 * p[1] * p[2], e2 = p[8] * p[8] - 2 * p[6] * p[7];
    const double pen = 1000.0 * ((e1 > 0 ? e1 : 0.0) + (e2 > 0 ? e2 : 0.0));
    out[b] = bad ? 1e10 : pairwise_sum(rel2, M) / M + pen;
    free(rel2);
  }
}

int oracle_threads(void) {
#ifdef _OPENMP
  extern int omp_get_max_threads(void);
  return omp_get_max_threads();
#else
  return 1;
#endif
}
