// TEST-ONLY host emulation of the kernel arithmetic in option-pricing-ffn-lbfgs_b200/csrc/dhj_math.cuh.
//
// Compiles the very same scalar functions the CUDA kernels call (DHJ_HD expands to `inline` under
// g++) and walks k in the kernel's order (k = lane + 32*i, per-lane partial sums, xor-butterfly across
// the 32 lanes), so the CPU test-suite can check the kernel's algebra against the golden fixtures
// without a GPU.  It differs from the device only in libm (glibc vs CUDA sin/cos/exp/log/atan2).
// This is NOT a CPU fallback: nothing in the product package loads it; tests/test_host_emu.py builds it
// on the fly with g++ into tests/host_emu/_build/.
#include "dhj_math.cuh"

using namespace dhj;

static double price_one(const Params& m, double S0, double K, double T, double r, double q, int is_call,
                        int N, double L, double* ab) {
  SetConsts s = make_set_consts(m, r, q);
  double a0, b0;
  truncation_range(m, T, r, L, &a0, &b0);
  StrikeConsts sc = make_strike_consts(K, S0);
  double a = py_min(a0, sc.x - 0.1), b = py_max(b0, sc.x + 0.1);
  if (ab) { ab[0] = a; ab[1] = b; }
  PassConsts p = make_pass_consts(s, a, b, T);
  double lane_sum[32];
  for (int lane = 0; lane < 32; ++lane) {
    double acc = 0.0;
    for (int k = lane; k < N; k += 32) {
      KTerm t = make_kterm(s, p, k);
      acc += payoff_term(t, p, sc, S0, is_call != 0, k);
    }
    lane_sum[lane] = acc;
  }
  for (int off = 16; off >= 1; off >>= 1)
    for (int lane = 0; lane < 32; ++lane) lane_sum[lane] += lane_sum[lane ^ off] * ((lane & off) ? 0.0 : 1.0);
  return fm::exp_(-r * T) * lane_sum[0];
}

extern "C" {

void emu_price_list(const double* params, const double* S0, int s0_stride, const double* strike,
                    int strike_stride, const double* maturity, const int* is_call, double r, double q,
                    long P, int M, int N, double L, double* out, double* ab) {
  for (long p = 0; p < P; ++p) {
    Params m = load_params(params + 13 * p);
    for (int o = 0; o < M; ++o)
      out[p * M + o] = price_one(m, S0[p * s0_stride], strike[p * strike_stride + o], maturity[o], r, q,
                                 is_call[o], N, L, ab ? ab + 2 * (p * M + o) : nullptr);
  }
}

void emu_loss(const double* x, long B, double S0, double r, const double* strike, const double* maturity,
              const int* is_call, const double* market, int M, int N, double* out) {
  for (long b = 0; b < B; ++b) {
    Params m = transform_params(x + 13 * b);
    bool bad = false;
    double sq = 0.0;
    for (int o = 0; o < M; ++o) {
      double pr = price_one(m, S0, strike[o], maturity[o], r, 0.0, is_call[o], N, 10.0, nullptr);
      if (!(pr > 0.0) || isinf(pr)) bad = true;
      double rel = (pr - market[o]) / market[o];
      sq += rel * rel;
    }
    out[b] = bad ? 1e10 : sq / (double)M + feller_penalty(m);
  }
}

void emu_cf(const double* params, double r, double q, double tau, const double* us, int n, double* re, double* im) {
  // full CF value phi(u) (without the e^{-iua} rotation) for checking against cf_values.npz:
  // uses a = 0 so that G = Re(phi); the imaginary part comes from a second pass rotated by pi/2.
  Params m = load_params(params);
  SetConsts s = make_set_consts(m, r, q);
  for (int i = 0; i < n; ++i) {
    // w chosen so that u_k = (k*pi)/w reproduces us[i] for k = 1 is not exact; evaluate the factors directly
    double u = us[i];
    double xr, xi;
    cf_exponent(s, u, tau, s.lam * tau, &xr, &xi);
    re[i] = exp(xr) * cos(xi);
    im[i] = exp(xr) * sin(xi);
  }
}
}
