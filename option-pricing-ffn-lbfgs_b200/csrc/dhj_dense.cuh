// dhj_dense.cuh — pricing kernel for slices with many strikes (the 200 x 20 dense surface of C3, or any
// option list with more than 8 strikes per maturity): one block per (parameter set, maturity slice).
//
//   * thread 0 prepares the item (per-set constants, truncation range, pass constants);
//   * strikes are processed in chunks of 256: all 128 threads prepare strike constants
//     (K, log(K/S0), exp(.), binding/call flags, rotation step cos/sin(theta_j));
//   * per 128-wide k-block every thread evaluates the CF at its k and leaves the strike-independent
//     coefficients (P, Q, R) in its warp's stage; each lane then contracts ONE STRIKE against the warp's 32
//     coefficients by plane rotations (segment_sums, one exact sincos at the segment start) — 7 FMAs per
//     (strike, k) instead of a sincos + 15 flops;
//   * strikes whose +-0.1 widening binds get their own pass (own (a,b), own CF), one strike at a time.
#pragma once
#include "dhj_batch.cuh"

namespace dhj {

constexpr int kDenseChunk = 256;
#ifndef DHJ_DENSE_MINB
#define DHJ_DENSE_MINB 6
#endif

struct DenseSmem {
  SetConsts set;
  PassConsts pass, extra_pass;
  double a0, b0, S0, disc;
  double extra_cth, extra_sth;
  CoefStage stage[kBatchWarps];
  fm::Tables ltab;
  double K[kDenseChunk], x[kDenseChunk], ex[kDenseChunk], cth[kDenseChunk], sth[kDenseChunk];
  double partial[kBatchWarps][kDenseChunk];
  unsigned char call[kDenseChunk], bind[kDenseChunk];
  int n_bind;
};

// one pass over the cosine terms for the strikes [0, cnt) of the chunk with bind flag == want_bind_idx semantics:
//   single < 0 : all strikes with bind == 0 (lane per strike, rounds of 32)
//   single >= 0: only strike `single` (lane 0 of each warp)
__device__ __forceinline__ void dense_pass(DenseSmem& sm, const PassConsts& pc, const double* __restrict__ cth,
                                           const double* __restrict__ sth, int cnt, int single, int n_cos, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  CoefStage& st = sm.stage[warp];
#pragma unroll 1
  for (int k0 = 0; k0 < n_cos; k0 += kBatchThreads) {
    const int k = k0 + tid;
    KCoef c;
    c.P = c.Q = c.R = c.a1 = c.a2 = c.g0 = 0.0;
    if (k < n_cos) c = make_kcoef(make_kterm(sm.set, pc, k, &sm.ltab), pc, k);
    __syncwarp();
    st.PQ[lane] = make_double2(c.P, c.Q); st.R[lane] = c.R;
    const double A1 = warp_sum(c.a1), A2 = warp_sum(c.a2), A3 = warp_sum(c.P);
    const double g0 = __shfl_sync(kFullMask, c.g0, 0);
    __syncwarp();
    // frequency of the warp's first term
    const double kpi = (double)(k - lane) * kPi;
    const double q0 = kpi * pc.rw;
    const double u0 = fma(fma(-pc.w, q0, kpi), pc.rw, q0);
    const int t_lo = (single < 0) ? lane : single + lane * kDenseChunk;     // lane 0 only when single
    const int t_hi = (single < 0) ? cnt : single + 1;
#pragma unroll 1
    for (int t = t_lo; t < t_hi; t += 32) {
      if (single < 0 && sm.bind[t]) continue;
      double sn, cs, spq, sr;
      fm::sincos_(u0 * (sm.x[t] - pc.a), &sn, &cs);
      const int ti = (single < 0) ? t : 0;
      segment_sums<32>(st.PQ, reinterpret_cast<const Pair*>(st.R), cs, sn, cth[ti], sth[ti], &spq, &sr);
      const double val = (sm.K[t] * sr - (sm.S0 * sm.ex[t]) * spq) +
                         strike_const_part(sm.call[t] != 0, sm.S0, sm.K[t], sm.x[t], pc, A1, A2, A3, g0);
      sm.partial[warp][t] += val;
    }
  }
}

}  // namespace dhj
