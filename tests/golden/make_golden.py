#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the LIVE, UNMODIFIED reference.

Runs only in the build container (it needs /root/reference, which does not exist on the GPU
box).  Nothing in tests/, bench.py or smoke() imports this module; they read the fixtures it
wrote.  Usage (from the repo root):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py [--only NAME ...] [--slow]

Fixtures (all float64, written with numpy.savez so values round-trip bit-exactly):

  known_answers.npz    SURVEY §8c literals recomputed: test-suite parameters on the 15-option grid
                       (tests/test_suite.py:197-201), demo call/put (double_heston.py:202-276)
  prices_grid15.npz    150 parameter sets in the generator's ranges (synthetic_generator.py:75-89)
                       x 5 strikes x 3 maturities, per-set spot, N=128, plus (a,b)
  dense_surface.npz    4 parameter sets x 20 strikes x 5 maturities at N=256, and an edge variant
                       (short maturities / far strikes) where the +-0.1 widening binds
  edge_cases.npz       Appendix-B style single prices (puts, q, degenerate parameters, N sweep)
  cf_values.npz        characteristic_function(u, tau) samples (double_heston.py:48-97)
  cf_complex.npz       the same at complex phi
  loss_cases.npz       compute_loss at 64 x vectors incl. Feller-active and sentinel cases, and
                       the 14-evaluation forward-difference stencil scipy uses (lbfgs_calibrator.py:118-177)
  initial_guess.npz    get_initial_guess(0/1/2) under np.random.seed(0) (lbfgs_calibrator.py:179-234)
  generator_seed42.npz generate_synthetic_calibrations(20) under np.random.seed(42)
  calib_ensemble.npz   (--slow) final losses of the reference from ulp-perturbed copies of start 1's x0
  calib_noisy.npz      (--slow) reference optimiser trajectories (every x, every loss) on 4 noisy generator-style
                       markets, starts 0 and 2
  calib_trajectory.npz (--slow) every x the reference optimiser visits for the C1 market,
                       np.random.seed(0), starts 0..2, with losses, nit, messages
"""
import argparse
import contextlib
import io
import os
import sys
import tempfile
import time

import numpy as np

REF = "/root/reference"
for sub in ("src/models", "src/calibration", "src/data"):
    sys.path.insert(0, os.path.join(REF, sub))
sys.dont_write_bytecode = True

from double_heston import DoubleHeston  # noqa: E402  (the reference)
from lbfgs_calibrator import DoubleHestonJumpCalibrator  # noqa: E402
import synthetic_generator as ref_gen  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# order of the 13 model parameters everywhere in this repo = calibrator x-vector order
# (lbfgs_calibrator.py:53-57) = DoubleHeston ctor order (double_heston.py:26-27)
PNAMES = ["v01", "kappa1", "theta1", "sigma1", "rho1", "v02", "kappa2", "theta2", "sigma2",
          "rho2", "lambda_j", "mu_j", "sigma_j"]
GEN_RANGES = np.array([
    (0.025, 0.080), (1.5, 4.5), (0.025, 0.065), (0.20, 0.50), (-0.85, -0.40),
    (0.020, 0.070), (0.30, 1.20), (0.025, 0.070), (0.10, 0.35), (-0.70, -0.20),
    (0.05, 0.25), (-0.08, -0.01), (0.03, 0.12)])
TEST_SUITE_PARAMS = np.array([0.04, 2.0, 0.04, 0.3, -0.5, 0.04, 1.5, 0.04, 0.2, -0.3, 0.1, 0.0, 0.1])
DEMO_PARAMS = np.array([0.04, 2.0, 0.04, 0.3, -0.5, 0.04, 1.5, 0.04, 0.2, -0.3, 0.5, -0.05, 0.10])


def ref_price(p, S0, K, T, r, opt="call", q=0.0, N=128, want_ab=False):
    dh = DoubleHeston(S0, K, T, r, *[float(v) for v in p], option_type=opt, q=q)
    with np.errstate(all="ignore"):
        price = dh.pricing(N=N)
        if want_ab:
            a, b = dh.truncationRange()
            return float(price), float(a), float(b)
    return float(price)


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez(path, **arrays)
    print(f"wrote {path}  ({os.path.getsize(path)} bytes)")


def make_known_answers():
    Ks = np.array([90.0, 95.0, 100.0, 105.0, 110.0])
    Ts = np.array([0.25, 0.5, 1.0])
    prices = np.zeros((3, 5))
    ab = np.zeros((3, 2))
    for i, T in enumerate(Ts):
        for j, K in enumerate(Ks):
            prices[i, j], a, b = ref_price(TEST_SUITE_PARAMS, 100.0, K, T, 0.05, want_ab=True)
            if j == 2:
                ab[i] = (a, b)
    demo_call = ref_price(DEMO_PARAMS, 100, 100, 1.0, 0.05, "C")
    demo_put = ref_price(DEMO_PARAMS, 100, 100, 1.0, 0.05, "P")
    save("known_answers.npz", params=TEST_SUITE_PARAMS, strikes=Ks, maturities=Ts, S0=100.0, r=0.05,
         prices=prices, ab=ab, demo_params=DEMO_PARAMS, demo_call=demo_call, demo_put=demo_put)


def make_prices_grid15():
    rng = np.random.default_rng(20260101)
    P = 150
    params = rng.uniform(GEN_RANGES[:, 0], GEN_RANGES[:, 1], size=(P, 13))
    spots = np.where(np.arange(P) % 3 == 0, 100.0, rng.uniform(80.0, 125.0, size=P))
    k_rel = np.array([90.0, 95.0, 100.0, 105.0, 110.0])
    Ts = np.array([0.25, 0.5, 1.0])
    r = 0.03
    prices = np.zeros((P, 3, 5))
    ab = np.zeros((P, 3, 5, 2))
    for p in range(P):
        for i, T in enumerate(Ts):
            for j, kr in enumerate(k_rel):
                K = kr * spots[p] / 100.0  # synthetic_generator.py:125
                prices[p, i, j], ab[p, i, j, 0], ab[p, i, j, 1] = ref_price(
                    params[p], spots[p], K, T, r, want_ab=True)
    save("prices_grid15.npz", params=params, spots=spots, k_rel=k_rel, maturities=Ts, r=r,
         prices=prices, ab=ab)


def make_dense_surface():
    rng = np.random.default_rng(20260102)
    P = 4
    params = rng.uniform(GEN_RANGES[:, 0], GEN_RANGES[:, 1], size=(P, 13))
    out = {}
    for tag, Ks, Ts in (
        ("main", np.linspace(80, 120, 200)[::10], np.linspace(0.25, 2.0, 20)[::4]),
        ("edge", np.linspace(50, 150, 200)[::10], np.array([0.02, 0.05, 0.1, 0.25, 2.0])),
    ):
        prices = np.zeros((P, len(Ts), len(Ks)))
        ab = np.zeros((P, len(Ts), len(Ks), 2))
        for p in range(P):
            for i, T in enumerate(Ts):
                for j, K in enumerate(Ks):
                    prices[p, i, j], ab[p, i, j, 0], ab[p, i, j, 1] = ref_price(
                        params[p], 100.0, K, T, 0.03, N=256, want_ab=True)
        out[f"{tag}_strikes"] = Ks
        out[f"{tag}_maturities"] = Ts
        out[f"{tag}_prices"] = prices
        out[f"{tag}_ab"] = ab
    save("dense_surface.npz", params=params, S0=100.0, r=0.03, N=256, **out)


def make_edge_cases():
    base = TEST_SUITE_PARAMS
    cases = []  # (params, S0, K, T, r, q, is_call, N)

    def add(p=base, S0=100.0, K=100.0, T=1.0, r=0.05, q=0.0, call=1, N=128):
        cases.append((np.array(p, dtype=float), S0, K, T, r, q, call, N))

    add()
    add(call=0)
    add(q=0.02)
    add(q=0.02, call=0)
    p = base.copy(); p[10:] = 0.0; add(p)                      # lambda=mu=sigma_j=0
    for s in (1e-6, 1e-3, 5.0):
        p = base.copy(); p[3] = s; add(p)
    for rho in (-1.0, 1.0, 0.0):
        p = base.copy(); p[4] = rho; add(p)
    p = base.copy(); p[1] = 1e-8; add(p)                       # kappa tiny -> NaN
    p = base.copy(); p[1] = np.inf; add(p)                     # kappa inf  -> NaN
    p = base.copy(); p[[0, 2, 5, 7]] = 1e-12; add(p)
    p = base.copy(); p[10] = 50.0; add(p)
    p = base.copy(); p[12] = 2.0; add(p)
    p = base.copy(); p[11] = -3.0; add(p)
    for T in (1e-4, 1e-2, 0.02, 5.0, 30.0):
        add(T=T)
        add(T=T, call=0)
    for K in (1.0, 10.0, 60.0, 80.0, 120.0, 150.0, 1000.0, 10000.0):
        add(K=K)
        add(K=K, call=0)
    add(T=0.02, K=60.0)
    add(T=0.02, K=140.0, call=0)
    for N in (1, 2, 3, 16, 31, 32, 33, 64, 100, 127, 129, 200, 256, 512):
        add(N=N)
        add(N=N, call=0, T=0.5, K=95.0)
    add(S0=3127.5, K=3000.0, T=0.75, r=0.01)
    add(p=DEMO_PARAMS)
    add(p=DEMO_PARAMS, call=0)

    n = len(cases)
    params = np.stack([c[0] for c in cases])
    meta = np.array([[c[1], c[2], c[3], c[4], c[5], c[6], c[7]] for c in cases], dtype=float)
    prices = np.zeros(n)
    ab = np.zeros((n, 2))
    for i, (p, S0, K, T, r, q, call, N) in enumerate(cases):
        prices[i], ab[i, 0], ab[i, 1] = ref_price(p, S0, K, T, r, "call" if call else "put", q, int(N),
                                                  want_ab=True)
    save("edge_cases.npz", params=params, meta=meta, prices=prices, ab=ab,
         meta_cols=np.array(["S0", "K", "T", "r", "q", "is_call", "N"]))


def make_cf_values():
    rng = np.random.default_rng(20260103)
    P = 6
    params = np.vstack([TEST_SUITE_PARAMS, DEMO_PARAMS,
                        rng.uniform(GEN_RANGES[:, 0], GEN_RANGES[:, 1], size=(P - 2, 13))])
    taus = np.array([0.02, 0.25, 1.0, 2.0])
    us = np.concatenate([[0.0, 1e-8, 0.1], np.linspace(0.5, 140.0, 29)])
    r, q = 0.03, 0.01
    cf = np.zeros((P, len(taus), len(us)), dtype=np.complex128)
    for p in range(P):
        dh = DoubleHeston(100, 100, 1.0, r, *[float(v) for v in params[p]], option_type="C", q=q)
        for i, tau in enumerate(taus):
            for j, u in enumerate(us):
                cf[p, i, j] = dh.characteristic_function(float(u), float(tau))
    save("cf_values.npz", params=params, taus=taus, us=us, r=r, q=q, cf=cf)


def make_cf_complex():
    """characteristic_function at COMPLEX phi (the reference's arithmetic is complex throughout and accepts it)."""
    rng = np.random.default_rng(20260106)
    params = np.vstack([TEST_SUITE_PARAMS, DEMO_PARAMS, rng.uniform(GEN_RANGES[:, 0], GEN_RANGES[:, 1], size=(2, 13))])
    taus = np.array([0.25, 1.0])
    us = np.concatenate([[0.0 - 1.0j, 0.5 - 0.5j, 1e-8 + 0.0j], rng.uniform(0, 40, 17) + 1j * rng.uniform(-1.0, 0.5, 17)])
    r, q = 0.03, 0.01
    cf = np.zeros((len(params), len(taus), len(us)), dtype=np.complex128)
    for p in range(len(params)):
        dh = DoubleHeston(100, 100, 1.0, r, *[float(v) for v in params[p]], option_type="C", q=q)
        for i, tau in enumerate(taus):
            for j, u in enumerate(us):
                cf[p, i, j] = dh.characteristic_function(complex(u), float(tau))
    save("cf_complex.npz", params=params, taus=taus, us=us, r=r, q=q, cf=cf)


def c1_market():
    """The noise-free 15-option market of tests/test_suite.py:274-302."""
    opts = []
    for T in [0.25, 0.5, 1.0]:
        for K in [90, 95, 100, 105, 110]:
            price = ref_price(TEST_SUITE_PARAMS, 100.0, K, T, 0.05)
            opts.append({"strike": K, "maturity": T, "price": price, "option_type": "call"})
    return 100.0, 0.05, opts


def market_arrays(opts):
    return (np.array([o["strike"] for o in opts], dtype=float),
            np.array([o["maturity"] for o in opts], dtype=float),
            np.array([1.0 if o["option_type"].upper()[0] == "C" else 0.0 for o in opts]),
            np.array([o["price"] for o in opts], dtype=float))


def make_loss_cases():
    spot, r, opts = c1_market()
    # a second, noisy, mixed call/put market on a ragged (non-grid) option list
    rng = np.random.default_rng(20260104)
    opts2 = []
    p2 = rng.uniform(GEN_RANGES[:, 0], GEN_RANGES[:, 1])
    for T, K, typ in [(0.25, 95, "put"), (0.25, 100, "call"), (0.25, 105, "call"), (0.5, 90, "put"),
                      (0.5, 100, "call"), (0.75, 100, "P"), (1.0, 80, "Call"), (1.0, 100, "call"),
                      (1.0, 120, "call"), (1.5, 110, "x"), (0.25, 110, "c")]:
        price = ref_price(p2, 103.5, K, T, 0.02, typ)
        opts2.append({"strike": float(K), "maturity": T, "price": price * (1 + 0.02 * rng.standard_normal()),
                      "option_type": typ})
    out = {}
    for tag, (s, rr, oo) in (("c1", (spot, r, opts)), ("ragged", (103.5, 0.02, opts2))):
        cal = DoubleHestonJumpCalibrator(s, rr, oo)
        xs = [cal.get_initial_guess(0), cal.get_initial_guess(2)]
        x_true = cal.inverse_transform_params(dict(zip(cal.param_names, TEST_SUITE_PARAMS)))
        xs.append(x_true)
        for _ in range(24):
            xs.append(x_true + 0.15 * rng.standard_normal(13))
        for _ in range(3):                                   # Feller violated (sigma large)
            x = x_true + 0.1 * rng.standard_normal(13); x[3] += 1.5; xs.append(x)
        x = x_true.copy(); x[1] = -40.0; xs.append(x)        # kappa ~ 4e-18 -> NaN -> 1e10
        x = x_true.copy(); x[1] = 800.0; xs.append(x)        # kappa = inf  -> 1e10
        x = x_true.copy(); x[[0, 2, 5, 7]] = -60.0; xs.append(x)   # tiny variances, finite loss
        xs = np.array(xs)
        losses = np.array([float(cal.compute_loss(x)) for x in xs])
        # forward-difference stencil exactly as scipy does it (scipy/optimize/_numdiff.py
        # _dense_difference, '2-point', abs_step=1e-8): x_i + h, dx = (x_i + h) - x_i
        h = 1e-8
        n_fd = 6
        fd_f = np.zeros((n_fd, 14))
        fd_g = np.zeros((n_fd, 13))
        for c in range(n_fd):
            x = xs[c]
            fd_f[c, 0] = cal.compute_loss(x)
            for i in range(13):
                xp = x.copy(); xp[i] = x[i] + h
                fd_f[c, 1 + i] = cal.compute_loss(xp)
                fd_g[c, i] = (fd_f[c, 1 + i] - fd_f[c, 0]) / (xp[i] - x[i])
        # cross-check against scipy itself
        from scipy.optimize._numdiff import approx_derivative
        g_scipy = approx_derivative(cal.compute_loss, xs[0], method="2-point", abs_step=h)
        assert np.array_equal(g_scipy, fd_g[0]), (g_scipy, fd_g[0])
        K, T, C, M = market_arrays(oo)
        out.update({f"{tag}_spot": s, f"{tag}_r": rr, f"{tag}_strike": K, f"{tag}_maturity": T,
                    f"{tag}_is_call": C, f"{tag}_market": M, f"{tag}_x": xs, f"{tag}_loss": losses,
                    f"{tag}_fd_f": fd_f, f"{tag}_fd_g": fd_g})
    save("loss_cases.npz", **out)


def make_initial_guess():
    spot, r, opts = c1_market()
    cal = DoubleHestonJumpCalibrator(spot, r, opts)
    np.random.seed(0)
    g0 = cal.get_initial_guess(0)
    g1 = cal.get_initial_guess(1)
    g2 = cal.get_initial_guess(2)
    g1b = cal.get_initial_guess(1)
    K, T, C, M = market_arrays(opts)
    save("initial_guess.npz", g0=g0, g1=g1, g2=g2, g1_second_draw=g1b, strike=K, maturity=T, market=M,
         spot=spot, r=r, f_g0=float(cal.compute_loss(g0)), f_g1=float(cal.compute_loss(g1)),
         f_g2=float(cal.compute_loss(g2)))


def make_generator():
    np.random.seed(42)
    with tempfile.TemporaryDirectory() as td, contextlib.redirect_stdout(io.StringIO()):
        res = ref_gen.generate_synthetic_calibrations(20, os.path.join(td, "x.pkl"))
    names = ["v1_0", "kappa1", "theta1", "sigma1", "rho1", "v2_0", "kappa2", "theta2", "sigma2", "rho2",
             "lambda_j", "mu_j", "sigma_j"]
    save("generator_seed42.npz",
         params=np.array([[c.parameters[n] for n in names] for c in res]),
         spots=np.array([c.spot for c in res]),
         model_prices=np.array([c.model_prices for c in res]),
         market_prices=np.array([c.market_prices for c in res]),
         losses=np.array([c.final_loss for c in res]),
         strikes=np.array([[o["strike"] for o in c.market_options] for c in res]),
         maturities=np.array([[o["maturity"] for o in c.market_options] for c in res]),
         dates=np.array([c.date for c in res]),
         messages=np.array([c.message for c in res]))


def make_calib_trajectory():
    from scipy.optimize import minimize
    spot, r, opts = c1_market()
    np.random.seed(0)
    out = {}
    t00 = time.time()
    for start in range(3):
        cal = DoubleHestonJumpCalibrator(spot, r, opts)
        x0 = cal.get_initial_guess(start % 3)
        xs, fs = [], []

        def f(x, cal=cal, xs=xs, fs=fs):
            v = cal.compute_loss(x)
            xs.append(np.array(x, dtype=float)); fs.append(float(v))
            return v
        t0 = time.time()
        res = minimize(fun=f, x0=x0, method="L-BFGS-B",
                       options={"maxiter": 300, "ftol": 1e-9, "gtol": 1e-6, "disp": False})
        print(f"start {start}: nit={res.nit} nfev={res.nfev} fun={res.fun!r} msg={res.message} "
              f"{time.time() - t0:.1f}s", flush=True)
        out.update({f"s{start}_x0": x0, f"s{start}_xs": np.array(xs), f"s{start}_fs": np.array(fs),
                    f"s{start}_x": res.x, f"s{start}_fun": float(res.fun), f"s{start}_nit": int(res.nit),
                    f"s{start}_success": bool(res.success), f"s{start}_message": str(res.message)})
    K, T, C, M = market_arrays(opts)
    save("calib_trajectory.npz", spot=spot, r=r, strike=K, maturity=T, is_call=C, market=M,
         wall_seconds=time.time() - t00, **out)


def _ensemble_member(k):
    """Reference optimiser from start 1's x0 scaled by (1 + k * 2^-52): SURVEY H1's noise-floor probe."""
    from scipy.optimize import minimize
    spot, r, opts = c1_market()
    np.random.seed(0)
    cal = DoubleHestonJumpCalibrator(spot, r, opts)
    x0 = cal.get_initial_guess(1) * (1.0 + k * 2.0 ** -52)
    res = minimize(fun=cal.compute_loss, x0=x0, method="L-BFGS-B",
                   options={"maxiter": 300, "ftol": 1e-9, "gtol": 1e-6, "disp": False})
    return k, float(res.fun), int(res.nit), int(res.nfev), str(res.message)


def make_calib_ensemble():
    """Final losses of the REFERENCE for 1-ulp-level perturbations of the start-1 initial point (C1 market,
    np.random.seed(0)): the spread the reference shows against itself, which any other FP64 implementation
    (different libm, different summation order) necessarily falls into."""
    import multiprocessing as mp
    ks = [-3, -2, -1, 0, 1, 2, 3, 5]
    with mp.get_context("fork").Pool(len(ks)) as pool:
        out = pool.map(_ensemble_member, ks)
    for row in out:
        print(row, flush=True)
    save("calib_ensemble.npz", k=np.array([o[0] for o in out]), fun=np.array([o[1] for o in out]),
         nit=np.array([o[2] for o in out]), nfev=np.array([o[3] for o in out]),
         message=np.array([o[4] for o in out]))


def _noisy_member(job):
    """One reference optimiser run (jac=None, scipy's own forward differences) on a NOISY generator-style market,
    logging every x it evaluates."""
    from scipy.optimize import minimize
    m, start, spot, r, opts = job
    cal = DoubleHestonJumpCalibrator(spot, r, opts)
    x0 = cal.get_initial_guess(start)                      # 0 and 2 only: no RNG involved
    xs, fs = [], []

    def f(x):
        v = cal.compute_loss(x)
        xs.append(np.array(x, dtype=float)); fs.append(float(v))
        return v
    t0 = time.time()
    res = minimize(fun=f, x0=x0, method="L-BFGS-B",
                   options={"maxiter": 300, "ftol": 1e-9, "gtol": 1e-6, "disp": False})
    print(f"market {m} start {start}: nit={res.nit} nfev={res.nfev} fun={res.fun!r} msg={res.message} "
          f"{time.time() - t0:.1f}s", flush=True)
    return m, start, x0, np.array(xs), np.array(fs), res.x, float(res.fun), int(res.nit), str(res.message)


def make_calib_noisy():
    """(--slow) Reference trajectories on 4 noisy markets built by the generator's recipe (uniform parameters in
    its ranges, K = K_rel * spot / 100, r = 0.03, market = price + N(0, 0.02) * price: synthetic_generator.py:98-142),
    starts 0 (literature) and 2 (ATM-implied): every x the optimiser evaluates, with its loss."""
    import multiprocessing as mp
    rng = np.random.default_rng(20260105)
    n_markets = 4
    jobs, out = [], {}
    for m in range(n_markets):
        p = rng.uniform(GEN_RANGES[:, 0], GEN_RANGES[:, 1])
        spot = float(rng.uniform(85.0, 120.0))
        opts = []
        for T in [0.25, 0.5, 1.0]:
            for kr in [90, 95, 100, 105, 110]:
                K = kr * spot / 100.0
                price = ref_price(p, spot, K, T, 0.03)
                opts.append({"strike": K, "maturity": T, "price": price + rng.normal(0, 0.02) * price,
                             "option_type": "call"})
        K, T, C, M = market_arrays(opts)
        out.update({f"m{m}_spot": spot, f"m{m}_params": p, f"m{m}_strike": K, f"m{m}_maturity": T, f"m{m}_market": M})
        for start in (0, 2):
            jobs.append((m, start, spot, 0.03, opts))
    with mp.get_context("fork").Pool(min(8, len(jobs))) as pool:
        done = pool.map(_noisy_member, jobs, chunksize=1)
    for m, start, x0, xs, fs, x, fun, nit, msg in done:
        tag = f"m{m}_s{start}"
        out.update({f"{tag}_x0": x0, f"{tag}_xs": xs, f"{tag}_fs": fs, f"{tag}_x": x, f"{tag}_fun": fun,
                    f"{tag}_nit": nit, f"{tag}_message": msg})
    save("calib_noisy.npz", r=0.03, n_markets=n_markets, **out)


MAKERS = {
    "known_answers": make_known_answers,
    "prices_grid15": make_prices_grid15,
    "dense_surface": make_dense_surface,
    "edge_cases": make_edge_cases,
    "cf_values": make_cf_values,
    "cf_complex": make_cf_complex,
    "loss_cases": make_loss_cases,
    "initial_guess": make_initial_guess,
    "generator": make_generator,
}
SLOW = {"calib_trajectory": make_calib_trajectory, "calib_ensemble": make_calib_ensemble,
        "calib_noisy": make_calib_noisy}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    ap.add_argument("--slow", action="store_true", help="also run the ~7 min optimiser trajectory capture")
    args = ap.parse_args()
    todo = dict(MAKERS)
    if args.slow:
        todo.update(SLOW)
    if args.only:
        allm = {**MAKERS, **SLOW}
        todo = {k: allm[k] for k in args.only}
    for name, fn in todo.items():
        t0 = time.time()
        fn()
        print(f"  {name}: {time.time() - t0:.1f}s", flush=True)
