#!/usr/bin/env python3
"""Developer probe: where the wall time of dhj.calibrate_many goes (host optimiser vs loss launches).
usage: python scripts/profile_calibrate_many.py [n_markets]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "option-pricing-ffn-lbfgs_b200")
sys.path.insert(0, PKG)
import dhj  # noqa: E402
from dhj.calibrate_many import initial_guesses  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ctx = dhj.default_context()
R = np.array([(0.025, 0.080), (1.5, 4.5), (0.025, 0.065), (0.20, 0.50), (-0.85, -0.40), (0.020, 0.070), (0.30, 1.20),
              (0.025, 0.070), (0.10, 0.35), (-0.70, -0.20), (0.05, 0.25), (-0.08, -0.01), (0.03, 0.12)])
rng = np.random.default_rng(7)
params = rng.uniform(R[:, 0], R[:, 1], size=(n, 13))
spots = 100.0 * np.exp(0.05 * rng.standard_normal(n))
K = np.array([90.0, 95.0, 100.0, 105.0, 110.0]); T = np.array([0.25, 0.5, 1.0])
strikes = np.tile(K[None, :] * spots[:, None] / 100.0, (1, 3)); mats = np.repeat(T, 5)
model = ctx.price_grid(params, spots, K, T, 0.03, scale_by_spot=True).reshape(n, 15)
market_p = model * (1 + 0.02 * rng.standard_normal((n, 15)))
for rep in range(3):
    np.random.seed(1)
    t = {"guess": 0.0, "ask": 0.0, "loss": 0.0, "tell": 0.0}
    t0 = time.perf_counter()
    x0 = initial_guesses(spots, strikes, mats, market_p, 3).reshape(n * 3, 13)
    t["guess"] = time.perf_counter() - t0
    mk = ctx.market(spots, 0.03, strikes, mats, np.ones(15, dtype=np.int32), market_p)
    sm = np.repeat(np.arange(n, dtype=np.int32), 3)
    opt = dhj.BatchLBFGS(x0, maxiter=300, ftol=1e-9, gtol=1e-6)
    rounds = 0
    active, per_round = [], []
    while True:
        a = time.perf_counter(); idx, x = opt.ask(); b = time.perf_counter()
        if idx.size == 0:
            break
        f, g = mk.loss_fd(x, 1e-8, market_index=sm[idx]); c = time.perf_counter()
        opt.tell(f, g); d = time.perf_counter()
        t["ask"] += b - a; t["loss"] += c - b; t["tell"] += d - c
        rounds += 1; active.append(idx.size); per_round.append(c - b)
    total = time.perf_counter() - t0
    opt.close(); mk.close()
    print(f"rep {rep}: total {total:.3f} s, rounds {rounds}, state-evaluations {sum(active)}: " +
          ", ".join(f"{k} {v:.3f}" for k, v in t.items()) + f"  (OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS')})")
    if rep == 2:
        act, tr = np.array(active), np.array(per_round)
        ideal = act * 210 / 1.23e9
        print(f"   loss time {tr.sum():.3f} s vs ideal kernel time at 1.23e9 prices/s {ideal.sum():.3f} s")
        for lo_b, hi_b in ((0, 64), (64, 1024), (1024, 8192), (8192, 20000), (20000, 10**9)):
            sel = (act >= lo_b) & (act < hi_b)
            if sel.any():
                print(f"   rounds with {lo_b}..{hi_b} active states: {sel.sum():4d} rounds, {act[sel].sum():8d} state-evals, "
                      f"loss time {tr[sel].sum() * 1e3:7.1f} ms (ideal {ideal[sel].sum() * 1e3:7.1f} ms), "
                      f"mean {tr[sel].mean() * 1e6:7.0f} us/round")
