// dhj_draws.cpp — host side of the synthetic generator: the sequential draw stream
// (/root/reference/src/data/synthetic_generator.py:98-142) at native speed.
//
// The reference draws, per sample and from NumPy's GLOBAL legacy RandomState (MT19937): 13 uniforms (parameter
// ranges in dict order), one normal for the spot return when i > 0, then 15 normals of price noise; parameters
// are AR(1)-smoothed and the spot is a random walk.  The stream is inherently sequential (one generator, draws
// interleaved), and a Python loop over it costs ~15 us per sample — minutes at the sample counts the GPU prices
// in milliseconds.  This file restates exactly what NumPy's legacy code path computes so that a seeded run
// reproduces the reference bit for bit:
//   * MT19937 (mt19937_gen / mt19937_next), double = ((a >> 5) * 2^26 + (b >> 6)) / 2^53;
//   * uniform(lo, hi) = lo + (hi - lo) * double;
//   * legacy_gauss: Marsaglia polar method, the second variate of each pair is cached (has_gauss / gauss);
//     normal(loc, scale) = loc + scale * gauss.
// The caller passes the generator state in (np.random.get_state()) and gets the advanced state back
// (np.random.set_state()), so native and Python draws can be mixed freely.
// Build note: no -mfma / -ffast-math for this translation unit (a fused lo + range * x would change bits).
#include <cmath>
#include <cstdint>
#include <cstring>

#include "dhj.h"

namespace {

constexpr int kN = 624, kM = 397;

struct Mt {
  uint32_t key[kN];
  int pos;
  int has_gauss;
  double gauss;

  void refill() {
    constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;
    int i;
    uint32_t y;
    for (i = 0; i < kN - kM; ++i) {
      y = (key[i] & kUpper) | (key[i + 1] & kLower);
      key[i] = key[i + kM] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
    }
    for (; i < kN - 1; ++i) {
      y = (key[i] & kUpper) | (key[i + 1] & kLower);
      key[i] = key[i + (kM - kN)] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
    }
    y = (key[kN - 1] & kUpper) | (key[0] & kLower);
    key[kN - 1] = key[kM - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
    pos = 0;
  }
  uint32_t next32() {
    if (pos >= kN) refill();
    uint32_t y = key[pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  double next_double() {
    const int32_t a = (int32_t)(next32() >> 5), b = (int32_t)(next32() >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
  }
  double next_gauss() {
    if (has_gauss) {
      const double t = gauss;
      has_gauss = 0;
      gauss = 0.0;
      return t;
    }
    double x1, x2, r2;
    do {
      x1 = 2.0 * next_double() - 1.0;
      x2 = 2.0 * next_double() - 1.0;
      r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    const double f = std::sqrt(-2.0 * std::log(r2) / r2);
    gauss = f * x1;
    has_gauss = 1;
    return f * x2;
  }
};

}  // namespace

extern "C" int dhj_generator_draws(const uint32_t* mt_key, int32_t mt_pos, int32_t has_gauss, double cached_gauss,
                                   int64_t n, int32_t n_params, const double* lo, const double* hi, double persistence,
                                   double spot0, double ret_mean, double ret_sd, double noise_sd, int32_t n_noise,
                                   double* params, double* spots, double* noise, uint32_t* out_key, int32_t* out_pos,
                                   int32_t* out_has_gauss, double* out_cached_gauss) {
  if (!mt_key || !lo || !hi || !out_key || !out_pos || !out_has_gauss || !out_cached_gauss || n < 0 || n_params < 0 ||
      n_noise < 0 || mt_pos < 0 || mt_pos > kN || (n > 0 && (!params || !spots || !noise)))
    return DHJ_ERR_ARG;
  Mt g;
  std::memcpy(g.key, mt_key, sizeof(g.key));
  g.pos = mt_pos;
  g.has_gauss = has_gauss != 0;
  g.gauss = cached_gauss;
  const double fresh_weight = 1.0 - persistence;            // Python: (1 - persistence), rounded once
  for (int64_t i = 0; i < n; ++i) {
    double* p = params + i * n_params;
    for (int j = 0; j < n_params; ++j) {
      const double range = hi[j] - lo[j];
      const double scaled = range * g.next_double();
      p[j] = lo[j] + scaled;
    }
    if (i > 0) {
      const double* prev = p - n_params;
      for (int j = 0; j < n_params; ++j) {
        const double a = persistence * prev[j], b = fresh_weight * p[j];
        p[j] = a + b;
      }
      const double step = ret_sd * g.next_gauss();
      const double growth = 1.0 + (ret_mean + step);
      spots[i] = spots[i - 1] * growth;
    } else {
      spots[i] = spot0;
    }
    double* z = noise + i * n_noise;
    for (int j = 0; j < n_noise; ++j) {
      const double step = noise_sd * g.next_gauss();
      z[j] = 0.0 + step;
    }
  }
  std::memcpy(out_key, g.key, sizeof(g.key));
  *out_pos = g.pos;
  *out_has_gauss = g.has_gauss;
  *out_cached_gauss = g.gauss;
  return DHJ_OK;
}
