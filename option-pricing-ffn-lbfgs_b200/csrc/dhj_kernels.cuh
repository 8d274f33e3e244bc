// dhj_kernels.cuh — the CUDA kernels of libdhj.so (sm_100a).
//
//   k_price      one warp per (parameter set, maturity slice): prices -> HBM        (K1 / K3 of SURVEY §2)
//   k_loss       one block per loss evaluation: exp/tanh transform, prices of every market option,
//                relative-MSE + Feller penalty + 1e10 sentinel; in FD mode the 14 evaluations of one
//                optimiser state are 14 blocks and the last one to finish assembles the gradient (K2)
//   k_fp64_peak  DFMA-chain probe for the FP64 roofline denominator
#pragma once
#include "dhj_engine.cuh"
#include "dhj_batch.cuh"

namespace dhj {

constexpr int kWarpsPerBlock = 4;
constexpr int kThreadsPerBlock = 32 * kWarpsPerBlock;
constexpr int kFdPoints = kNumParams + 1;      // f(x) and 13 forward points
constexpr double kSentinel = 1e10;             // lbfgs_calibrator.py:152-153

struct PriceArgs {
  const double* params;      // [P][13] model parameters, or unconstrained x when transform != 0
  const double* S0;          // spot table
  long long s0_stride;       // 0 = scalar
  const int* row_index;      // optional [P]: row of S0 / strike tables used by set p (default p)
  long long P;
  int transform;
  double* out;               // [P][M]
};

__global__ void __launch_bounds__(kThreadsPerBlock, 4) k_price(SliceView v, PriceArgs a) {
  __shared__ WarpSmem ws_all[kWarpsPerBlock];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpSmem& ws = ws_all[warp];
  const long long n_items = a.P * (long long)v.n_slices;
  const long long stride = (long long)gridDim.x * kWarpsPerBlock;
  for (long long item = (long long)blockIdx.x * kWarpsPerBlock + warp; item < n_items; item += stride) {
    const long long p = item / v.n_slices;
    const int s = (int)(item - p * v.n_slices);
    const long long row = a.row_index ? (long long)a.row_index[p] : p;
    const double* pp = a.params + kNumParams * p;
    const Params m = a.transform ? transform_params(pp) : load_params(pp);
    const double S0 = a.S0[row * a.s0_stride];
    const double* strike_row = v.strike + row * v.strike_stride;
    double* out_row = a.out + p * (long long)v.n_options;
    price_slice(ws, m, v, s, S0, strike_row, lane,
                [&](int o, double price) { out_row[v.pos[o]] = price; });
  }
}

// Throughput variant of k_price for slices of <= 8 strikes: see dhj_batch.cuh.
__global__ void __launch_bounds__(kBatchThreads, DHJ_BATCH_MINB) k_price_batch(SliceView v, PriceArgs a) {
  __shared__ BatchSmem sm;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long n_items = a.P * (long long)v.n_slices;
  const long long n_batches = (n_items + kBatchItems - 1) / kBatchItems;
  for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
    const long long base = batch * kBatchItems;
    const int cnt_items = (int)min((long long)kBatchItems, n_items - base);
    // ---- phase 1: one thread per item ----------------------------------------------------------
    if (tid < cnt_items) {
      const long long item = base + tid;
      const long long p = item / v.n_slices;
      const int s = (int)(item - p * v.n_slices);
      const long long row = a.row_index ? (long long)a.row_index[p] : p;
      prepare_item(sm.items[tid], v, a.params + kNumParams * p, a.transform != 0, a.S0[row * a.s0_stride],
                   v.strike + row * v.strike_stride, s, p * (long long)v.n_options);
    }
    __syncthreads();
    // ---- phase 2: one thread per cosine index ----------------------------------------------------
#pragma unroll 1
    for (int i = 0; i < cnt_items; ++i) {
      const ItemRec& it = sm.items[i];
      double* warp_partial = sm.partial[i][warp];
      if (lane < kBatchMaxStrikes) warp_partial[lane] = 0.0;
      __syncwarp();
      const unsigned reg_mask = it.valid_mask & ~it.bind_mask;
      if (reg_mask) contract_pass(it, it.pass, it.cth, it.sth, reg_mask, v.n_cos, tid, sm.stage[warp], warp_partial);
      unsigned todo = it.valid_mask & it.bind_mask;          // rare: strikes with their own (a, b)
      while (todo) {
        const int j = __ffs(todo) - 1;
        todo &= todo - 1;
        __syncthreads();
        if (tid == 0) {
          sm.extra_pass = make_pass_consts(it.set, py_min(it.a0, it.x[j] - 0.1), py_max(it.b0, it.x[j] + 0.1),
                                           it.pass.T);
          fm::sincos_(u_one(sm.extra_pass) * (it.x[j] - sm.extra_pass.a), &sm.extra_sth, &sm.extra_cth);
        }
        __syncthreads();
        // the task code indexes cth/sth by strike: point it at the single extra entry
        contract_pass(it, sm.extra_pass, &sm.extra_cth - j, &sm.extra_sth - j, 1u << j, v.n_cos, tid, sm.stage[warp],
                      warp_partial);
      }
    }
    __syncthreads();
    // ---- phase 3: add the warps' partials, discount, store -----------------------------------------
    for (int t = tid; t < cnt_items * kBatchMaxStrikes; t += kBatchThreads) {
      const int i = t / kBatchMaxStrikes, j = t - i * kBatchMaxStrikes;
      const ItemRec& it = sm.items[i];
      if (it.valid_mask & (1u << j)) {
        const double* q = sm.partial[i][0] + j;
        const double sum = ((q[0] + q[kBatchMaxStrikes]) + q[2 * kBatchMaxStrikes]) + q[3 * kBatchMaxStrikes];
        a.out[it.out_row + v.pos[it.o_lo + j]] = it.disc * sum;
      }
    }
    __syncthreads();
  }
}

struct LossArgs {
  const double* x;           // [B][13] unconstrained
  const int* market_index;   // optional [B] (FD mode: [C])
  const double* S0;          // [n_markets]
  const double* market;      // [n_markets][M] caller order
  int fd;                    // 0: one block per x ; 1: 14 blocks per x
  double h;
  double* f_all;             // [B] (fd: [C][14] scratch)
  double* fg;                // fd: [C][14] = f, g[13]
  unsigned int* counters;    // fd: [C], zero on entry, zero on exit
};

__global__ void __launch_bounds__(kThreadsPerBlock, 4) k_loss(SliceView v, LossArgs a) {
  __shared__ WarpSmem ws_all[kWarpsPerBlock];
  __shared__ double red_sq[kWarpsPerBlock];
  __shared__ int red_bad[kWarpsPerBlock];
  __shared__ int is_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_warps = blockDim.x >> 5;
  const long long b = blockIdx.x;
  const long long c = a.fd ? b / kFdPoints : b;
  const int var = a.fd ? (int)(b - c * kFdPoints) : 0;

  double xv[kNumParams];
#pragma unroll
  for (int i = 0; i < kNumParams; ++i) xv[i] = a.x[kNumParams * c + i];
  // scipy's forward point: x_i + h  (_numdiff.py _dense_difference, h = abs_step)
#pragma unroll
  for (int i = 0; i < kNumParams; ++i)
    if (var == i + 1) xv[i] = xv[i] + a.h;
  const Params m = transform_params(xv);
  const long long mi = a.market_index ? a.market_index[c] : 0;
  const double S0 = a.S0[mi];
  const double* strike_row = v.strike + mi * v.strike_stride;
  const double* market_row = a.market + mi * (long long)v.n_options;

  double sq = 0.0;
  int bad = 0;
  for (int s = warp; s < v.n_slices; s += n_warps) {
    price_slice(ws_all[warp], m, v, s, S0, strike_row, lane, [&](int o, double price) {
      // lbfgs_calibrator.py:152: isnan or isinf or <= 0
      if (!(price > 0.0) || isinf(price)) bad = 1;
      const double mk = market_row[v.pos[o]];
      const double rel = (price - mk) / mk;               // :163
      sq += rel * rel;
    });
  }
  sq = warp_sum(sq);
  bad = __any_sync(kFullMask, bad);
  if (lane == 0) { red_sq[warp] = sq; red_bad[warp] = bad; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    int any_bad = 0;
    for (int w = 0; w < n_warps; ++w) { tot += red_sq[w]; any_bad |= red_bad[w]; }
    const double loss = any_bad ? kSentinel : tot / (double)v.n_options + feller_penalty(m);   // :164-169
    a.f_all[b] = loss;
    if (a.fd) {
      __threadfence();
      const unsigned prev = atomicAdd(&a.counters[c], 1u);
      is_last = (prev == (unsigned)(kFdPoints - 1));
    }
  }
  if (!a.fd) return;
  __syncthreads();
  if (is_last) {
    __threadfence();
    const double* f = a.f_all + kFdPoints * c;
    if (threadIdx.x < kNumParams) {
      const int i = threadIdx.x;
      const double xi = a.x[kNumParams * c + i];
      const double dx = (xi + a.h) - xi;
      a.fg[kFdPoints * c + 1 + i] = (__ldcg(f + 1 + i) - __ldcg(f)) / dx;
    } else if (threadIdx.x == kNumParams) {
      a.fg[kFdPoints * c] = __ldcg(f);
      a.counters[c] = 0u;
    }
  }
}

// ---- the remaining public methods of DoubleHeston, one thread per element ----------------------
// characteristic_function(phi, tau) for n frequencies of one parameter set (double_heston.py:48-97)
__global__ void k_cf(const double* __restrict__ params, double r, double q, double tau,
                     const double* __restrict__ u_in, int n, double* __restrict__ out_re,
                     double* __restrict__ out_im) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Params m = load_params(params);
  const SetConsts s = make_set_consts(m, r, q);
  const double u = u_in[i];
  double xr, xi;
  cf_exponent(s, u, tau, s.lam * tau, &xr, &xi);
  double sn, cs;
  fm::sincos_(xi, &sn, &cs);
  const double mag = fm::exp_(xr);
  out_re[i] = mag * cs;
  out_im[i] = mag * sn;
}

// truncationRange(L) for P parameter sets x M (S0, K, T) triples (double_heston.py:100-139)
__global__ void k_truncation_range(const double* __restrict__ params, long long P, const double* __restrict__ S0,
                                   long long s0_stride, const double* __restrict__ strike,
                                   const double* __restrict__ maturity, int M, double r, double L,
                                   double* __restrict__ out_ab) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * M) return;
  const long long p = i / M;
  const int o = (int)(i - p * M);
  const Params m = load_params(params + kNumParams * p);
  double a0, b0;
  truncation_range(m, maturity[o], r, L, &a0, &b0);
  const double x = fm::log_ratio(strike[o], S0[p * s0_stride]);
  out_ab[2 * i] = py_min(a0, x - 0.1);
  out_ab[2 * i + 1] = py_max(b0, x + 0.1);
}

// chi_k, psi_k for n values of k (double_heston.py:141-158)
__global__ void k_chi_psi(const int* __restrict__ k_in, int n, double c, double d, double a, double b,
                          double* __restrict__ chi, double* __restrict__ psi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = k_in[i];
  const double ed = fm::exp_(d), ec = fm::exp_(c);
  if (k == 0) { chi[i] = ed - ec; psi[i] = d - c; return; }
  const double u = fm::div((double)k * kPi, b - a);
  double sd, cd, sc, cc;
  fm::sincos_(u * (d - a), &sd, &cd);
  fm::sincos_(u * (c - a), &sc, &cc);
  chi[i] = fm::rcp(1.0 + u * u) * (((cd * ed - cc * ec) + (u * sd) * ed) - (u * sc) * ec);
  psi[i] = fm::rcp(u) * (sd - sc);
}

// 8 independent FMA chains per thread; 2 flop per FMA
constexpr int kPeakChains = 8;
constexpr int kPeakUnroll = 8;
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double x, double y) {
  double acc[kPeakChains];
#pragma unroll
  for (int j = 0; j < kPeakChains; ++j) acc[j] = 1.0 + 1e-3 * (double)(threadIdx.x + j);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < kPeakUnroll; ++r) {
#pragma unroll
      for (int j = 0; j < kPeakChains; ++j) acc[j] = fma(acc[j], x, y);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < kPeakChains; ++j) s += acc[j];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace dhj
