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

static double butterfly_total(double* v) {        // xor-butterfly over 32 lanes, as warp_sum does
  for (int off = 16; off >= 1; off >>= 1) {
    double t[32];
    for (int l = 0; l < 32; ++l) t[l] = v[l] + v[l ^ off];
    for (int l = 0; l < 32; ++l) v[l] = t[l];
  }
  return v[0];
}

static double u_of(const PassConsts& p, int k) {
  const double kpi = (double)k * kPi;
  const double q0 = kpi * p.rw;
  return fma(fma(-p.w, q0, kpi), p.rw, q0);
}

// one option, walked the way k_price_batch does it: blocks of 128 k, 4 warps of 32 lanes, strike-independent
// coefficients, four butterfly sums, 8-term rotation segments, partials per warp added in warp order
static double price_one(const Params& m, double S0, double K, double T, double r, double q, int is_call,
                        int N, double L, double* ab) {
  SetConsts s = make_set_consts(m, r, q);
  double a0, b0;
  truncation_range(m, T, r, L, &a0, &b0);
  StrikeConsts sc = make_strike_consts(K, S0);
  double a = py_min(a0, sc.x - 0.1), b = py_max(b0, sc.x + 0.1);
  if (ab) { ab[0] = a; ab[1] = b; }
  PassConsts p = make_pass_consts(s, a, b, T);
  double sth, cth;
  fm::sincos_(u_of(p, 1) * (sc.x - p.a), &sth, &cth);
  double partial[4] = {0, 0, 0, 0};
  for (int k0 = 0; k0 < N; k0 += 128) {
    for (int w = 0; w < 4; ++w) {
      Pair PQ[32];
      alignas(16) double R[32];
      double a1[32], a2[32], a3[32], g0[32];
      for (int lane = 0; lane < 32; ++lane) {
        const int k = k0 + 32 * w + lane;
        KCoef c; c.P = c.Q = c.R = c.a1 = c.a2 = c.g0 = 0.0;
        if (k < N) c = make_kcoef(make_kterm(s, p, k, &fm::kTables), p, k);
        PQ[lane].x = c.P; PQ[lane].y = c.Q; R[lane] = c.R; a1[lane] = c.a1; a2[lane] = c.a2; a3[lane] = c.P; g0[lane] = c.g0;
      }
      const double A1 = butterfly_total(a1), A2 = butterfly_total(a2), A3 = butterfly_total(a3), G0 = butterfly_total(g0);
      double val[4];
      for (int sg = 0; sg < 4; ++sg) {
        const int kstart = k0 + 32 * w + 8 * sg;
        double sn, cs, spq, sr;
        fm::sincos_(u_of(p, kstart) * (sc.x - p.a), &sn, &cs);
        segment_sums<8>(PQ + 8 * sg, reinterpret_cast<const Pair*>(R + 8 * sg), cs, sn, cth, sth, &spq, &sr);
        val[sg] = sc.K * sr - (S0 * sc.ex) * spq;
        if (sg == 0) val[sg] += strike_const_part(is_call != 0, S0, sc.K, sc.x, p, A1, A2, A3, G0);
      }
      partial[w] += (val[0] + val[1]) + (val[2] + val[3]);
    }
  }
  return fm::exp_(-r * T) * (((partial[0] + partial[1]) + partial[2]) + partial[3]);
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
    cf_exponent(s, u, tau, s.lam * tau, &fm::kTables, &xr, &xi);
    re[i] = exp(xr) * cos(xi);
    im[i] = exp(xr) * sin(xi);
  }
}
}
