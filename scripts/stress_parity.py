#!/usr/bin/env python3
"""Developer probe: wide-range randomised parity of the GPU pricer against the C restatement of the reference
(oracle/cos_oracle.c, itself pinned to the golden fixtures) — parameters far outside the generator's ranges, short
and long maturities, deep in/out-of-the-money strikes, puts and calls, several N.  Errors are judged relative to
the price where it is not tiny and relative to the scale of the sum otherwise: a call's payoff coefficients carry e^b, so for wild parameters
(b = c1 + 10 sqrt(c2) up to 10 or more) the reference's own sum cancels catastrophically and both implementations
return rounding noise of size ~ S0 e^b eps N; the error is therefore also reported in those units.
usage (GPU box): python scripts/stress_parity.py [n_sets]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "option-pricing-ffn-lbfgs_b200"))
sys.path.insert(0, ROOT)
import dhj  # noqa: E402
from oracle import cos_oracle as O  # noqa: E402  (checker only)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
ctx = dhj.Context(0)
rng = np.random.default_rng(2026)
lo = np.array([0.005, 0.1, 0.005, 0.05, -0.99, 0.005, 0.1, 0.005, 0.05, -0.99, 0.0, -0.3, 0.01])
hi = np.array([0.25, 10.0, 0.25, 1.0, 0.5, 0.25, 10.0, 0.25, 1.0, 0.5, 2.0, 0.2, 0.4])
worst = {}
for N in (64, 128, 256):
    params = rng.uniform(lo, hi, size=(n, 13))
    spots = rng.uniform(20.0, 500.0, size=n)
    rel_k = np.array([0.5, 0.8, 0.95, 1.0, 1.05, 1.25, 2.0])
    T = np.array([0.02, 0.1, 0.5, 1.0, 2.5, 5.0])
    K = np.repeat(spots[:, None], T.size * rel_k.size, axis=1) * np.tile(rel_k, T.size)[None, :]
    mats = np.repeat(T, rel_k.size)
    for call in (1, 0):
        flags = np.full(mats.size, call)
        t0 = time.perf_counter()
        got = ctx.price_list(params, spots, K, mats, flags, 0.03, 0.01, N)
        t1 = time.perf_counter()
        want, ab = O.c_price_batch(params, spots, K, mats, flags, 0.03, 0.01, N, return_ab=True)
        t2 = time.perf_counter()
        # rounding noise of the reference's own call sum: terms of size S0 e^b, N of them
        noise = spots[:, None] * np.exp(np.maximum(ab[..., 1], 0.0)) * (N * 2.2e-16) if call else np.zeros_like(want)
        scale = np.maximum(np.abs(want), 1e-3 * spots[:, None])
        err = np.abs(got - want) / (scale + 30.0 * noise)
        bad_nan = np.isnan(got) != np.isnan(want)
        err = np.where(np.isnan(want), 0.0, err)
        i = np.unravel_index(np.argmax(err), err.shape)
        worst[(N, call)] = err.max()
        tame = np.isfinite(want) & (ab[..., 1] < 6.0)
        print(f"   b < 6 ({tame.mean():.1%} of the prices): max |got - want| / max(|want|, 1e-3 S0) = "
              f"{(np.abs(got - want) / scale)[tame].max():.2e}")
        print(f"N={N} {'calls' if call else 'puts '}: {got.size} prices, GPU {t1 - t0:.2f} s, oracle {t2 - t1:.1f} s, "
              f"max err {err.max():.2e} (median {np.median(err):.1e}) at set {i[0]} T={mats[i[1]]} K/S={K[i]/spots[i[0]]:.2f} "
              f"price {want[i]:.4g} b {ab[i][1]:.2f}; NaN mismatches {int(bad_nan.sum())}")
print("worst:", max(worst.values()))
