// dhj_engine.cuh — warp-level COS pass: one warp prices every option of one maturity slice of one
// parameter set.
//
// Work decomposition:
//   * a "slice" = the options of one market/grid that share a maturity T.  For one (parameter set,
//     slice) the truncation range (a0,b0), the frequencies u_k and the characteristic function do
//     not depend on the strike unless the +-0.1 widening of double_heston.py:135-137 binds, so the CF
//     is evaluated once per slice and contracted against every strike ("regular" strikes); a strike
//     for which the widening binds gets its own pass with its own (a,b) ("binding" strikes).
//   * cosine index k = kb*128 + 32*i + lane, i < 4: each lane owns 4 values of k per 128-wide
//     k-block; their KTerm (7 doubles each) are staged in shared memory, laid out [field][i][lane]
//     so that every access is a conflict-free 256-byte row.  N > 128 is handled by looping k-blocks
//     with the per-strike partial prices accumulated in shared memory.
//   * strikes are handled in chunks of 32: lane l prepares strike l of the chunk (K, log(K/S0),
//     exp(.), binding flag); the values are broadcast by shuffle when strike l is contracted against
//     the staged KTerms; a 5-step xor butterfly adds the 32 lanes' partial sums.
//   * per-warp constants (SetConsts, PassConsts) live in shared memory too (uniform, broadcast reads)
//     which keeps the register budget for the transcendental-heavy CF.
#pragma once
#include "dhj_math.cuh"

namespace dhj {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kKBlock = 128;           // cosine terms staged per k-block
constexpr int kKPerLane = kKBlock / 32;
constexpr int kMaxSliceStrikes = 256;  // strikes contracted per staged CF (larger slices loop)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(kFullMask, v, off);
  return v;
}

// Description of the option slices of a market / grid, device pointers (dhj_abi.cu: OptionBook)
struct SliceView {
  int n_slices;               // number of distinct maturities
  int n_options;              // M
  const double* slice_T;      // [n_slices]
  const int* slice_off;       // [n_slices+1] into the slice-ordered option arrays
  const double* strike;       // strikes in the CALLER's option order; row stride `strike_stride` per
                              // market / parameter set (0 = one shared row)
  const unsigned char* call;  // [M] slice-ordered, 1 = call
  const int* pos;             // [M] slice-ordered -> caller's option index
  long long strike_stride;
  int scale_by_spot;          // K = strike * S0 / 100   (synthetic_generator.py:125)
  int n_cos;                  // N
  double r, q, L;
};

// per-warp shared-memory workspace
struct WarpSmem {
  double G[kKBlock], u[kKBlock], inv1[kKBlock], invu[kKBlock], sb[kKBlock], t1[kKBlock], t3[kKBlock];
  double acc[kMaxSliceStrikes];
  SetConsts set;
  PassConsts pass;
};

// not inlined: called from the regular pass and from the binding-strike pass, and the CF it contains is
// by far the largest piece of code in the kernel (one copy keeps the hot loop inside the I-cache)
__device__ __noinline__ void stage_kblock(WarpSmem& ws, int k_base, int n_cos, int lane) {
#pragma unroll 1
  for (int i = 0; i < kKPerLane; ++i) {
    const int k = k_base + 32 * i + lane;
    if (k < n_cos) {
      const KTerm t = make_kterm(ws.set, ws.pass, k);
      const int slot = 32 * i + lane;
      ws.G[slot] = t.G; ws.u[slot] = t.u; ws.inv1[slot] = t.inv1; ws.invu[slot] = t.invu;
      ws.sb[slot] = t.sb; ws.t1[slot] = t.t1; ws.t3[slot] = t.t3;
    }
  }
}

__device__ __forceinline__ double strike_partial(const WarpSmem& ws, const StrikeConsts& sc, double S0,
                                                 bool is_call, int k_base, int n_cos, int lane) {
  double acc = 0.0;
#pragma unroll
  for (int i = 0; i < kKPerLane; ++i) {
    const int k = k_base + 32 * i + lane;
    if (k < n_cos) {
      const int slot = 32 * i + lane;
      KTerm t;
      t.G = ws.G[slot]; t.u = ws.u[slot]; t.inv1 = ws.inv1[slot]; t.invu = ws.invu[slot];
      t.sb = ws.sb[slot]; t.t1 = ws.t1[slot]; t.t3 = ws.t3[slot];
      acc += payoff_term(t, ws.pass, sc, S0, is_call, k);
    }
  }
  return acc;
}

// lane l prepares strike (c_lo + l) of the current chunk
struct ChunkStrikes {
  StrikeConsts mine;
  bool call, bind;
};

__device__ __forceinline__ ChunkStrikes prepare_chunk(const SliceView& v, const double* __restrict__ strike_row,
                                                      double S0, double a0, double b0, int c_lo, int cnt,
                                                      int lane) {
  ChunkStrikes c;
  c.mine.K = 1.0; c.mine.x = 0.0; c.mine.ex = 1.0;
  c.call = true; c.bind = false;
  if (lane < cnt) {
    double K = strike_row[v.pos[c_lo + lane]];
    if (v.scale_by_spot) K = K * S0 / 100.0;
    c.mine = make_strike_consts(K, S0);
    c.call = v.call[c_lo + lane] != 0;
    // the widening binds <=> Python's min/max (double_heston.py:136-137) would replace a0 / b0
    c.bind = ((c.mine.x - 0.1) < a0) || ((c.mine.x + 0.1) > b0);
  }
  return c;
}

__device__ __forceinline__ StrikeConsts bcast_strike(const StrikeConsts& mine, int j) {
  StrikeConsts s;
  s.K = __shfl_sync(kFullMask, mine.K, j);
  s.x = __shfl_sync(kFullMask, mine.x, j);
  s.ex = __shfl_sync(kFullMask, mine.ex, j);
  return s;
}

// Prices all options of slice `s_idx` for model `m`.  For every option, exactly one lane calls
// emit(option_index_in_slice_order, price).  All 32 lanes of the warp must call this together;
// `ws` is this warp's private workspace.
template <class Emit>
__device__ __forceinline__ void price_slice(WarpSmem& ws, const Params& m, const SliceView& v, int s_idx,
                                            double S0, const double* __restrict__ strike_row, int lane,
                                            Emit emit) {
  const double T = v.slice_T[s_idx];
  const int o_lo = v.slice_off[s_idx], o_hi = v.slice_off[s_idx + 1];
  const int n_cos = v.n_cos;
  const int n_kb = (n_cos + kKBlock - 1) / kKBlock;

  double a0, b0;
  truncation_range(m, T, v.r, v.L, &a0, &b0);
  const double disc = fm::exp_(-v.r * T);
  __syncwarp();
  if (lane == 0) ws.set = make_set_consts(m, v.r, v.q);
  __syncwarp();

  for (int g_lo = o_lo; g_lo < o_hi; g_lo += kMaxSliceStrikes) {
    const int g_hi = min(o_hi, g_lo + kMaxSliceStrikes);
    // the common case (<= 32 strikes in the slice) prepares its single chunk once
    const bool single = (g_hi - g_lo) <= 32;
    ChunkStrikes cs0;
    if (single) cs0 = prepare_chunk(v, strike_row, S0, a0, b0, g_lo, g_hi - g_lo, lane);
    // ---- phase A: strikes that share (a0, b0) -------------------------------------------------
    bool any_regular = false;
    for (int c_lo = g_lo; c_lo < g_hi; c_lo += 32) {
      const int cnt = min(32, g_hi - c_lo);
      const ChunkStrikes cs = single ? cs0 : prepare_chunk(v, strike_row, S0, a0, b0, c_lo, cnt, lane);
      const unsigned valid = (cnt == 32) ? kFullMask : ((1u << cnt) - 1u);
      if (valid & ~__ballot_sync(kFullMask, cs.bind)) any_regular = true;
      if (lane < cnt) ws.acc[c_lo - g_lo + lane] = 0.0;
    }
    if (any_regular) {
      __syncwarp();
      if (lane == 0) ws.pass = make_pass_consts(ws.set, a0, b0, T);
      __syncwarp();
      for (int b = 0; b < n_kb; ++b) {
        stage_kblock(ws, b * kKBlock, n_cos, lane);
        for (int c_lo = g_lo; c_lo < g_hi; c_lo += 32) {
          const int cnt = min(32, g_hi - c_lo);
          const ChunkStrikes cs = single ? cs0 : prepare_chunk(v, strike_row, S0, a0, b0, c_lo, cnt, lane);
          const unsigned valid = (cnt == 32) ? kFullMask : ((1u << cnt) - 1u);
          unsigned todo = valid & ~__ballot_sync(kFullMask, cs.bind);
          double mine_tot = 0.0;
          while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const StrikeConsts sj = bcast_strike(cs.mine, j);
            const bool cj = __shfl_sync(kFullMask, (int)cs.call, j) != 0;
            const double tot = warp_sum(strike_partial(ws, sj, S0, cj, b * kKBlock, n_cos, lane));
            if (lane == j) mine_tot = tot;
          }
          if (lane < cnt) ws.acc[c_lo - g_lo + lane] += mine_tot;
        }
      }
    }
    // ---- phase B: strikes whose widening binds: one pass each with their own (a, b) -----------
    for (int c_lo = g_lo; c_lo < g_hi; c_lo += 32) {
      const int cnt = min(32, g_hi - c_lo);
      const ChunkStrikes cs = single ? cs0 : prepare_chunk(v, strike_row, S0, a0, b0, c_lo, cnt, lane);
      const unsigned valid = (cnt == 32) ? kFullMask : ((1u << cnt) - 1u);
      unsigned todo = valid & __ballot_sync(kFullMask, cs.bind);
      while (todo) {
        const int j = __ffs(todo) - 1;
        todo &= todo - 1;
        const StrikeConsts sj = bcast_strike(cs.mine, j);
        const bool cj = __shfl_sync(kFullMask, (int)cs.call, j) != 0;
        __syncwarp();
        if (lane == 0) ws.pass = make_pass_consts(ws.set, py_min(a0, sj.x - 0.1), py_max(b0, sj.x + 0.1), T);
        __syncwarp();
        double tot = 0.0;
        for (int b = 0; b < n_kb; ++b) {
          stage_kblock(ws, b * kKBlock, n_cos, lane);
          tot += warp_sum(strike_partial(ws, sj, S0, cj, b * kKBlock, n_cos, lane));
        }
        if (lane == j) ws.acc[c_lo - g_lo + lane] = tot;
      }
    }
    // ---- phase C: hand the prices out ---------------------------------------------------------
    for (int c_lo = g_lo; c_lo < g_hi; c_lo += 32) {
      const int cnt = min(32, g_hi - c_lo);
      if (lane < cnt) emit(c_lo + lane, disc * ws.acc[c_lo - g_lo + lane]);
    }
  }
}

}  // namespace dhj
