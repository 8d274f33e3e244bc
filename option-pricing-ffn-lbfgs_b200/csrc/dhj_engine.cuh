// dhj_engine.cuh — shared pieces of the COS kernels: the option-slice view and warp reductions.
//
// A "slice" = the options of one market / grid that share a maturity T.  For one (parameter set, slice) the
// truncation range (a0,b0), the frequencies u_k and the characteristic function do not depend on the strike
// unless the +-0.1 widening of double_heston.py:135-137 binds, so the CF is evaluated once per slice and
// contracted against every strike ("regular" strikes); a strike for which the widening binds gets its own pass
// with its own (a,b) ("binding" strikes).  The kernels that do this are in dhj_batch.cuh (few strikes per slice)
// and dhj_dense.cuh (many).
#pragma once
#include "dhj_math.cuh"

namespace dhj {

constexpr unsigned kFullMask = 0xffffffffu;

// ---- checked build (-DDHJ_CHECKED: lib/libdhj_checked.so, tests/test_gpu_checked.py) ---------------------------------
// compute-sanitizer is closed on this pool, so the shared-memory protocol of the kernels is checked by the kernels
// themselves in a separate build: every stage of coefficients carries the epoch it was written in and the epoch up to
// which each lane has consumed it; a read before the matching write, a write before the readers are done, an index
// outside a shared array or an output index outside the caller's buffer increments a device counter that the host
// reads back (dhj_debug_checks).  The product build compiles the checks to nothing.
enum CheckCode { kChkReadBeforeWrite = 0, kChkWriteBeforeConsumed = 1, kChkSharedIndex = 2, kChkOutputIndex = 3,
                 kChkItemNotPrepared = 4, kChkCodes = 8 };
#if defined(DHJ_CHECKED) && defined(__CUDACC__)
__device__ unsigned long long g_check_fail[kChkCodes];
#define DHJ_CHECK(cond, code) do { if (!(cond)) atomicAdd(&dhj::g_check_fail[code], 1ULL); } while (0)
#else
#define DHJ_CHECK(cond, code) do { } while (0)
#endif

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

}  // namespace dhj
