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
#ifndef DHJ_BATCH_ITEMS
#define DHJ_BATCH_ITEMS 32
#endif
constexpr int kBatchItems = DHJ_BATCH_ITEMS;
constexpr int kBatchMaxStrikes = 8;
#ifndef DHJ_BATCH_MINB
#define DHJ_BATCH_MINB 7
#endif

struct ItemRec {
  SetConsts set;
  PassConsts pass;                 // regular pass: (a0, b0)
  double S0, disc;                 // (the regular range (a0, b0) is pass.a, pass.b)
  double K[kBatchMaxStrikes], x[kBatchMaxStrikes], ex[kBatchMaxStrikes];
  double cth[kBatchMaxStrikes], sth[kBatchMaxStrikes];      // cos / sin of theta_j = u_1 (x_j - a0)
  unsigned valid_mask, bind_mask, call_mask;
  int o_lo;                        // first option (slice order) of the slice
  long long out_row;               // p * M
};

struct CoefStage { double P[32], Q[32], R[32]; };   // one warp's k-block of strike-independent coefficients

struct BatchSmem {
  ItemRec items[kBatchItems];
  PassConsts extra_pass;           // pass constants of a binding strike
  double extra_cth, extra_sth;     // and its rotation step
  CoefStage stage[kBatchWarps];
  fm::Tables ltab;                 // fm::log_tab / exp_tab tables (per-lane index: shared memory, not the constant bank)
  double partial[kBatchItems][kBatchWarps][kBatchMaxStrikes];
};

struct PriceArgs;                  // dhj_kernels.cuh

// every block copies the log table into its shared memory once
__device__ __forceinline__ void load_log_table(fm::Tables* dst, int tid) {
  if (tid < 64) { dst->log[tid] = fm::kTables.log[tid]; dst->exp2[tid] = fm::kTables.exp2[tid]; }
  if (tid < 65) dst->atan64[tid] = fm::kTables.atan64[tid];
}

// u_1 = (1*pi)/(b-a), the rotation step's frequency (same correction step as make_kterm)
__device__ __forceinline__ double u_one(const PassConsts& p) {
  const double q0 = kPi * p.rw;
  return fma(fma(-p.w, q0, kPi), p.rw, q0);
}

// phase 1 for one item, executed by a single thread
__device__ __forceinline__ void prepare_item(ItemRec& rec, const SliceView& v, const Params& m, double S0,
                                             const double* __restrict__ strike_row, int s_idx, long long out_row) {
  rec.set = make_set_consts(m, v.r, v.q);
  const double T = v.slice_T[s_idx];
  double a0, b0;
  truncation_range(m, T, v.r, v.L, &a0, &b0);
  rec.pass = make_pass_consts(rec.set, a0, b0, T);
  rec.S0 = S0;
  rec.disc = fm::exp_(-v.r * T);
  const int o_lo = v.slice_off[s_idx], cnt = v.slice_off[s_idx + 1] - o_lo;
  unsigned bind = 0, call = 0;
#pragma unroll 1
  for (int j = 0; j < cnt; ++j) {
    double K = strike_row[v.pos[o_lo + j]];
    if (v.scale_by_spot) K = K * S0 / 100.0;
    const StrikeConsts sc = make_strike_consts(K, S0);
    rec.K[j] = sc.K; rec.x[j] = sc.x; rec.ex[j] = sc.ex;
    fm::sincos_(u_one(rec.pass) * (sc.x - a0), &rec.sth[j], &rec.cth[j]);
    if (((sc.x - 0.1) < a0) || ((sc.x + 0.1) > b0)) bind |= 1u << j;
    if (v.call[o_lo + j]) call |= 1u << j;
  }
  rec.valid_mask = (1u << cnt) - 1u;
  rec.bind_mask = bind; rec.call_mask = call;
  rec.o_lo = o_lo; rec.out_row = out_row;
}

// phase 2 body for one pass of one warp: CF at this lane's k -> strike-independent coefficients in the warp's
// stage; four shuffle reductions (A1, A2, A3, g0); then the rotation tasks: lane = (strike j, segment s) with
// 8-term segments, reduced over the 4 segments by two shuffles and accumulated into the warp's partial.
__device__ __forceinline__ void contract_pass(const ItemRec& it, const PassConsts& pc, const double* __restrict__ cth,
                                              const double* __restrict__ sth, unsigned mask, int n_cos, int tid,
                                              CoefStage& st, const fm::Tables* __restrict__ ltab,
                                              double* __restrict__ warp_partial) {
  constexpr int kSeg = 8, kNumSeg = 32 / kSeg;
  const int lane = tid & 31;
#pragma unroll 1
  for (int k0 = 0; k0 < n_cos; k0 += kBatchThreads) {
    const int k = k0 + tid;
    KCoef c;
    c.P = c.Q = c.R = c.a1 = c.a2 = c.g0 = 0.0;
    if (k < n_cos) c = make_kcoef(make_kterm(it.set, pc, k, ltab), pc, k);
    __syncwarp();
    st.P[lane] = c.P; st.Q[lane] = c.Q; st.R[lane] = c.R;
    // A1, A2 feed calls, A3 puts (uniform per pass); g0 is non-zero only in the lane that holds k = 0
    const bool any_call = (it.call_mask & mask) != 0, any_put = (~it.call_mask & mask) != 0;
    const double A1 = any_call ? warp_sum(c.a1) : 0.0, A2 = any_call ? warp_sum(c.a2) : 0.0;
    const double A3 = any_put ? warp_sum(c.P) : 0.0;
    const double g0 = __shfl_sync(kFullMask, c.g0, 0);
    __syncwarp();
    // task of this lane
    const int j = lane / kNumSeg, s = lane - j * kNumSeg;
    const bool active = (mask >> j) & 1u;
    double val = 0.0;
    if (active) {
      const int kstart = (k - lane) + s * kSeg;                    // absolute k of the segment's first term
      const double kpi = (double)kstart * kPi;
      const double q0 = kpi * pc.rw;
      const double u0 = fma(fma(-pc.w, q0, kpi), pc.rw, q0);
      double sn, cs;
      fm::sincos_(u0 * (it.x[j] - pc.a), &sn, &cs);
      double spq, sr;
      segment_sums(st.P + s * kSeg, st.Q + s * kSeg, st.R + s * kSeg, kSeg, cs, sn, cth[j], sth[j], &spq, &sr);
      val = it.K[j] * sr - (it.S0 * it.ex[j]) * spq;
      if (s == 0) val += strike_const_part((it.call_mask >> j) & 1u, it.S0, it.K[j], it.x[j], pc, A1, A2, A3, g0);
    }
    val += __shfl_xor_sync(kFullMask, val, 1);
    val += __shfl_xor_sync(kFullMask, val, 2);
    if (active && s == 0) warp_partial[j] += val;
  }
}

}  // namespace dhj
