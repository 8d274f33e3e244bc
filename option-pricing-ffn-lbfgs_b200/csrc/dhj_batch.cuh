// dhj_batch.cuh — throughput kernel for grids / option lists with at most 8 strikes per maturity slice
// (the 15-option grid of C2 / C4 / the generator / the calibrator's market).
//
// Decomposition: ONE THREAD PER COSINE INDEX k, a block of 128 threads walks a batch of
// 32 items (item = one (parameter set, maturity slice)):
//   phase 1  thread t < 32 prepares item t ALONE: parameters (optionally exp/tanh transform), per-set
//            constants, truncation range, pass constants, the slice's strikes (K, log(K/S0), exp(.),
//            binding flags) -> shared memory.  The prologue is therefore executed once per item by one
//            lane instead of redundantly by every lane of a warp (it was ~10 % of the warp-per-item
//            kernel's instructions).
//   phase 2  for each item, all 128 threads evaluate the CF at their own k (constants are broadcast
//            reads from shared memory; the KTerm lives in registers and is consumed at once by the
//            <= 8 strikes of the slice); each warp reduces its <= 8 partial prices by shuffles and
//            parks them in shared memory — no block barrier per item.
//   phase 3  after one barrier, thread t adds the four warps' partials of (item, strike) t and writes
//            the discounted price.
// A strike whose +-0.1 widening binds (double_heston.py:135-137) is contracted in an extra pass of
// phase 2 with its own (a,b): rare, uniform across the block.
#pragma once
#include "dhj_engine.cuh"

namespace dhj {

constexpr int kBatchThreads = 128;
constexpr int kBatchWarps = kBatchThreads / 32;
constexpr int kBatchItems = 32;
constexpr int kBatchMaxStrikes = 8;
#ifndef DHJ_BATCH_MINB
#define DHJ_BATCH_MINB 4
#endif

struct ItemRec {
  SetConsts set;
  PassConsts pass;                 // regular pass: (a0, b0)
  double a0, b0, S0, disc;
  double K[kBatchMaxStrikes], x[kBatchMaxStrikes], ex[kBatchMaxStrikes];
  unsigned valid_mask, bind_mask, call_mask;
  int o_lo;                        // first option (slice order) of the slice
  long long out_row;               // p * M
};

struct BatchSmem {
  ItemRec items[kBatchItems];
  PassConsts extra_pass;           // pass constants of a binding strike
  double partial[kBatchItems][kBatchWarps][kBatchMaxStrikes];
};

struct PriceArgs;                  // dhj_kernels.cuh

// phase 1 for one item, executed by a single thread
__device__ __forceinline__ void prepare_item(ItemRec& rec, const SliceView& v, const double* __restrict__ pp,
                                             bool transform, double S0, const double* __restrict__ strike_row,
                                             int s_idx, long long out_row) {
  const Params m = transform ? transform_params(pp) : load_params(pp);
  rec.set = make_set_consts(m, v.r, v.q);
  const double T = v.slice_T[s_idx];
  double a0, b0;
  truncation_range(m, T, v.r, v.L, &a0, &b0);
  rec.pass = make_pass_consts(rec.set, a0, b0, T);
  rec.a0 = a0; rec.b0 = b0; rec.S0 = S0;
  rec.disc = fm::exp_(-v.r * T);
  const int o_lo = v.slice_off[s_idx], cnt = v.slice_off[s_idx + 1] - o_lo;
  unsigned bind = 0, call = 0;
#pragma unroll 1
  for (int j = 0; j < cnt; ++j) {
    double K = strike_row[v.pos[o_lo + j]];
    if (v.scale_by_spot) K = K * S0 / 100.0;
    const StrikeConsts sc = make_strike_consts(K, S0);
    rec.K[j] = sc.K; rec.x[j] = sc.x; rec.ex[j] = sc.ex;
    if (((sc.x - 0.1) < a0) || ((sc.x + 0.1) > b0)) bind |= 1u << j;
    if (v.call[o_lo + j]) call |= 1u << j;
  }
  rec.valid_mask = (1u << cnt) - 1u;
  rec.bind_mask = bind; rec.call_mask = call;
  rec.o_lo = o_lo; rec.out_row = out_row;
}

// phase 2 body for one pass: this thread's k values against the strikes in `mask`; each strike's 32 lane terms
// are added by shuffles at once and accumulated into the warp's shared-memory partial, so no accumulator
// registers stay live across the (register-hungry) CF evaluation
__device__ __forceinline__ void contract_pass(const ItemRec& it, const PassConsts& pc, unsigned mask, int n_cos,
                                              int tid, double* __restrict__ warp_partial) {
  const int lane = tid & 31;
#pragma unroll 1
  for (int k0 = 0; k0 < n_cos; k0 += kBatchThreads) {
    const int k = k0 + tid;
    const bool live = k < n_cos;
    KTerm t;
    if (live) t = make_kterm(it.set, pc, k);
#pragma unroll 1
    for (int j = 0; j < kBatchMaxStrikes; ++j) {
      if (mask & (1u << j)) {
        StrikeConsts sc;
        sc.K = it.K[j]; sc.x = it.x[j]; sc.ex = it.ex[j];
        const double term = live ? payoff_term(t, pc, sc, it.S0, (it.call_mask >> j) & 1u, k) : 0.0;
        const double tot = warp_sum(term);
        if (lane == 0) warp_partial[j] += tot;
      }
    }
  }
}

}  // namespace dhj
