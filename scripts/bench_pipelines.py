#!/usr/bin/env python3
"""Developer probe: wall time of dhj.calibrate_many (10 000 markets x 3 starts) against the number of host pipelines.
usage (GPU box): python scripts/bench_pipelines.py [n_markets]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "option-pricing-ffn-lbfgs_b200"))
sys.path.insert(0, ROOT)
import dhj  # noqa: E402
from bench import GEN, C4_SEED, GRID_K, GRID_T  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ctx = dhj.default_context()
data = ctx.generate(C4_SEED, 0, n, **GEN)
spots, market = data["spots"], data["market"]
K = np.tile(np.array(GRID_K)[None, :] * spots[:, None] / 100.0, (1, 3)); T = np.repeat(np.array(GRID_T), 5)
np.random.seed(1)
x0 = dhj.initial_guesses(spots, K, T, market, 3)
ref = None
for pipes in (1, 2, 3, 4, 6, 8):
    dhj.calibrate_many(spots, 0.03, K, T, np.ones(15), market, maxiter=3, multi_start=3, x0=x0, pipelines=pipes)
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        res = dhj.calibrate_many(spots, 0.03, K, T, np.ones(15), market, maxiter=300, multi_start=3, x0=x0, pipelines=pipes)
        best = min(best, time.perf_counter() - t0)
    same = True if ref is None else bool(np.array_equal(ref, res["final_loss"]))
    ref = res["final_loss"] if ref is None else ref
    print(f"pipelines={pipes}: {best:.3f} s ({n / best:.0f} calibrations/s), rounds {res['rounds']}, same bits as pipelines=1: {same}, "
          f"loss {np.round(res['seconds_loss'], 3)}, ask {np.round(res['seconds_ask'], 3)}, tell {np.round(res['seconds_tell'], 3)}, "
          f"setup {np.round(res['seconds_setup'], 3)}, loop {np.round(res['seconds_loop'], 3)}")
