#!/usr/bin/env python3
"""Developer probe: kernel-only timing of k_price on the C2 / C3 shapes for the library named by
DHJ_LIBRARY (default: the in-tree build), plus a parity spot-check against the golden grid fixture.
Usage (on a GPU box): DHJ_LIBRARY=path/to/libdhj.so python scripts/quick_perf.py [c2|c3|both]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "option-pricing-ffn-lbfgs_b200"))
import torch  # noqa: E402

import dhj  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "both"
ctx = dhj.Context(0)
R = np.array([(0.025, 0.080), (1.5, 4.5), (0.025, 0.065), (0.20, 0.50), (-0.85, -0.40), (0.020, 0.070), (0.30, 1.20),
              (0.025, 0.070), (0.10, 0.35), (-0.70, -0.20), (0.05, 0.25), (-0.08, -0.01), (0.03, 0.12)])
g = np.load(os.path.join(ROOT, "tests", "golden", "prices_grid15.npz"))
got = ctx.price_grid(g["params"], g["spots"], g["k_rel"], g["maturities"], float(g["r"]), scale_by_spot=True)
err = np.abs(got - g["prices"]) / g["prices"]
tag = os.environ.get("DHJ_LIBRARY", "in-tree")
print(f"[{tag}] golden grid15: max rel err {err.max():.3e} median {np.median(err):.3e}")


def bench(P, strikes, mats, N, reps):
    rng = np.random.default_rng(0)
    params = torch.from_numpy(rng.uniform(R[:, 0], R[:, 1], size=(P, 13))).cuda()
    S0 = torch.full((1,), 100.0, dtype=torch.float64, device="cuda")
    out = torch.empty((P, len(mats), len(strikes)), dtype=torch.float64, device="cuda")

    def run():
        ctx.price_grid_dev(params.data_ptr(), P, S0.data_ptr(), 0, strikes, mats, 0.03, 0.0, N, 10.0, False, True,
                           out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n = P * len(mats) * len(strikes)
    return ms, n / ms * 1e3


if which in ("c2", "both", "all"):
    ms, pps = bench(1 << 20, np.array([90.0, 95, 100, 105, 110]), np.array([0.25, 0.5, 1.0]), 128, 5)
    print(f"[{tag}] C2 {ms:.2f} ms  {pps:.4g} prices/s")
if which in ("c3", "both", "all"):
    ms, pps = bench(1024, np.linspace(80, 120, 200), np.linspace(0.25, 2.0, 20), 256, 5)
    print(f"[{tag}] C3 {ms:.2f} ms  {pps:.4g} prices/s")
if which in ("loss", "all"):
    import time
    n = 10000
    rng = np.random.default_rng(3)
    params = rng.uniform(R[:, 0], R[:, 1], size=(n, 13))
    spots = 100 * np.exp(0.05 * rng.standard_normal(n))
    Ks = np.array([90.0, 95, 100, 105, 110]); Ts = np.array([0.25, 0.5, 1.0])
    strikes = np.tile(Ks[None, :] * spots[:, None] / 100, (1, 3)); mats = np.repeat(Ts, 5)
    market = ctx.price_grid(params, spots, Ks, Ts, 0.03, scale_by_spot=True).reshape(n, 15) * (1 + 0.02 * rng.standard_normal((n, 15)))
    mk = ctx.market(spots, 0.03, strikes, mats, np.ones(15), market)
    x = np.log(np.where(np.arange(13) == 11, 1.0, np.abs(params)))
    x[:, 4] = np.arctanh(params[:, 4]); x[:, 9] = np.arctanh(params[:, 9]); x[:, 11] = params[:, 11]
    x3 = np.repeat(x, 3, axis=0) + 0.05 * rng.standard_normal((3 * n, 13))
    idx = np.repeat(np.arange(n, dtype=np.int32), 3)
    mk.loss_fd(x3, 1e-8, idx)
    t0 = time.perf_counter()
    for _ in range(5):
        f, g = mk.loss_fd(x3, 1e-8, idx)
    dt = (time.perf_counter() - t0) / 5
    prices = 3 * n * 14 * 15
    print(f"[{tag}] loss_fd: {3 * n} optimiser states x 14 stencil points x 15 options = {prices} prices in {dt * 1e3:.2f} ms "
          f"(host call incl. copies) -> {prices / dt:.4g} prices/s, {3 * n / dt:.4g} f,g evaluations/s")
    x1 = x3[:3]
    mk1 = ctx.market(spots[0], 0.03, strikes[0], mats, np.ones(15), market[0])
    mk1.loss_fd(x1)
    t0 = time.perf_counter()
    for _ in range(200):
        mk1.loss_fd(x1)
    print(f"[{tag}] loss_fd latency, 3 states (one README calibration step): {(time.perf_counter() - t0) / 200 * 1e6:.1f} us per call")
