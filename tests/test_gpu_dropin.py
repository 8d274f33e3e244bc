"""GPU tests of the drop-in modules (the reference's Python API on top of libdhj.so)."""
import os
import pickle
import sys
import time

import numpy as np
import pytest

from conftest import PKG, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    for sub in ("models", "calibration", "data"):
        p = os.path.join(PKG, "src", sub)
        if p not in sys.path:
            sys.path.insert(0, p)
    import double_heston
    import lbfgs_calibrator
    import synthetic_generator
    return double_heston, lbfgs_calibrator, synthetic_generator


TS = dict(v01=0.04, kappa1=2.0, theta1=0.04, sigma1=0.3, rho1=-0.5, v02=0.04, kappa2=1.5, theta2=0.04, sigma2=0.2,
          rho2=-0.3, lambda_j=0.1, mu_j=0.0, sigma_j=0.1)


def test_double_heston_object(mods, golden):
    dh, _, _ = mods
    g = golden("known_answers.npz")
    for i, T in enumerate(g["maturities"]):
        for j, K in enumerate(g["strikes"]):
            price = dh.DoubleHeston(S0=100.0, K=float(K), T=float(T), r=0.05, option_type='call', **TS).pricing(N=128)
            assert isinstance(price, np.float64)
            assert rel_err(price, g["prices"][i, j]) <= 1e-10
    m = dh.DoubleHeston(100.0, 100.0, 1.0, 0.05, option_type='C', **TS)
    a, b = m.truncationRange()
    assert abs(a - g["ab"][2, 0]) <= 1e-14 and abs(b - g["ab"][2, 1]) <= 1e-14
    assert m.characteristic_function(0.0, 1.0) == 1.0
    assert np.iscomplexobj(m.characteristic_function(np.array([0.5, 1.0]), 1.0))
    assert m.psi_k(0, -0.1, b, a, b) == b + 0.1
    # first-letter rule (double_heston.py:172): every non-'C' string is a put
    call = [dh.DoubleHeston(100.0, 100.0, 1.0, 0.05, option_type=t, **TS).pricing() for t in ('call', 'C', 'c', 'Call')]
    put = [dh.DoubleHeston(100.0, 100.0, 1.0, 0.05, option_type=t, **TS).pricing() for t in ('put', 'P', 'x', 'european call')]
    assert len(set(call)) == 1 and len(set(put)) == 1
    assert rel_err(call[0], 13.545233402249403) <= 1e-10 and rel_err(put[0], 8.66817585183412) <= 1e-10
    assert rel_err(dh.DoubleHeston(100.0, 100.0, 1.0, 0.05, q=0.02, **TS).pricing(), 12.293326084471488) <= 1e-10


def test_reference_sanity_checks(mods):
    """The range/monotonicity checks of the reference's own suite (tests/test_suite.py:203-262)."""
    dh, _, _ = mods
    price = lambda K, T: dh.DoubleHeston(S0=100.0, K=K, T=T, r=0.05, option_type='call', **TS).pricing(N=128)
    assert 2.0 < price(100.0, 1.0) < 15.0
    ks = [price(K, 1.0) for K in (90, 95, 100, 105, 110)]
    assert np.sum(np.diff(ks) < 0) >= 3
    assert np.all(np.diff([price(100, T) for T in (0.25, 0.5, 1.0)]) > 0)
    assert all(np.isfinite(price(K, T)) for K, T in ((100, 0.25), (100, 2.0), (80, 1.0), (120, 1.0)))


def c1_calibrator(cal, g):
    opts = [{"strike": float(k), "maturity": float(t), "price": float(p), "option_type": "call"}
            for k, t, p in zip(g["strike"], g["maturity"], g["market"])]
    return cal.DoubleHestonJumpCalibrator(float(g["spot"]), float(g["r"]), opts)


def test_compute_loss_and_counters(mods, golden):
    _, cal, _ = mods
    g = golden("initial_guess.npz")
    c = c1_calibrator(cal, g)
    f0 = c.compute_loss(g["g0"])
    assert abs(f0 - float(g["f_g0"])) <= 1e-9 and c.n_calls == 1 and c.best_loss == f0
    f1 = c.compute_loss(g["g1"])
    assert abs(f1 - float(g["f_g1"])) <= 1e-9 and c.n_calls == 2 and c.best_loss == f0
    x = g["g0"].copy(); x[1] = 800.0
    assert c.compute_loss(x) == 1e10 and c.best_loss == f0          # sentinel does not touch best_loss
    f, grad = c.compute_loss_and_grad(g["g0"])
    assert f == f0 and grad.shape == (13,) and c.n_calls == 17


def test_calibrate_c1(mods, golden):
    """README configuration: 15 options, N=128, 13 parameters, multi_start=3, maxiter=300, np.random.seed(0)."""
    _, cal, _ = mods
    g = golden("calib_trajectory.npz")
    gi = golden("initial_guess.npz")
    c = c1_calibrator(cal, gi)
    c.calibrate(maxiter=2, multi_start=1)                            # warm-up (module import, first launches)
    np.random.seed(0)
    t0 = time.perf_counter()
    res = c.calibrate(maxiter=300, multi_start=3)
    wall = time.perf_counter() - t0
    ref_best = min(float(g[f"s{s}_fun"]) for s in range(3))
    print(f"calibrate(300,3): {wall:.3f} s wall, final_loss {res.final_loss:.6e} (reference best-of-3 {ref_best:.6e}), "
          f"nit {res.iterations}, '{res.message}'")
    assert isinstance(res, cal.CalibrationResult)
    assert wall < 1.0                                                # north-star: under 1 s
    assert res.final_loss * 100 < 1.0                                # the reference suite's own criterion (test 4.1)
    # L-BFGS-B on a forward-difference gradient with h = 1e-8 is chaotic at the 1e-16 level: the
    # REFERENCE started from x0 * (1 + k 2^-52), k = -3..5, ends between 3.5e-8 (47 iterations) and 7.97e-7
    # (17 iterations) — tests/golden/calib_ensemble.npz.  Any FP64 implementation with another libm falls
    # somewhere in that spread; the meaningful per-evaluation statement is test_trajectory_replay.
    ens = golden("calib_ensemble.npz")
    print("reference ulp-perturbation ensemble:", sorted(ens["fun"]), "nit", sorted(ens["nit"]))
    # two-sided: the best of three starts ends inside the reference's own spread (start 1 decides it; the per-start and
    # per-basin statements are tests/test_gpu_configs.py::test_final_loss_reproducible_starts / _ensemble_two_sided)
    assert 0.9 * ens["fun"].min() <= res.final_loss <= 1.01 * max(ref_best, ens["fun"].max())
    assert set(res.parameters) == set(c.param_names) and res.model_prices.shape == (15,)
    assert np.abs(res.model_prices - res.market_prices).max() / res.market_prices.max() < 0.01
    assert res.calibration_time <= wall and res.success in (True, False)
    # the reference suite's parameter-range check (tests/test_suite.py:327-344)
    boxes = {'v1_0': (0.001, 0.5), 'v2_0': (0.001, 0.5), 'kappa1': (0.1, 10.0), 'kappa2': (0.1, 10.0),
             'theta1': (0.001, 0.5), 'theta2': (0.001, 0.5), 'sigma1': (0.01, 2.0), 'sigma2': (0.01, 2.0),
             'rho1': (-1.0, 1.0), 'rho2': (-1.0, 1.0), 'lambda_j': (0.0, 5.0), 'sigma_j': (0.001, 1.0)}
    for name, (lo, hi) in boxes.items():
        assert lo <= res.parameters[name] <= hi, (name, res.parameters[name])
    # same starting points as the reference's run (global RNG order preserved)
    np.random.seed(0)
    assert np.array_equal(c.get_initial_guess(0), g["s0_x0"]) and np.array_equal(c.get_initial_guess(1), g["s1_x0"])


def test_lockstep_equals_sequential(mods, golden):
    _, cal, _ = mods
    gi = golden("initial_guess.npz")
    out = []
    for batched in (True, False):
        c = c1_calibrator(cal, gi)
        c.batch_starts = batched
        np.random.seed(0)
        t0 = time.perf_counter()
        r = c.calibrate(maxiter=40, multi_start=3)
        out.append((r.final_loss, r.iterations, tuple(r.parameters.values()), c.n_calls, c.best_loss))
        print("batched" if batched else "sequential", f"{time.perf_counter() - t0:.3f} s", r.final_loss, r.iterations)
    assert out[0] == out[1]                                          # batching starts does not change any trajectory


def test_generator_seed42(mods, golden, tmp_path, capsys):
    _, cal, gen = mods
    g = golden("generator_seed42.npz")
    np.random.seed(42)
    path = str(tmp_path / "synth.pkl")
    res = gen.generate_synthetic_calibrations(20, path)
    capsys.readouterr()
    assert len(res) == 20 and all(isinstance(r, cal.CalibrationResult) for r in res)
    model = np.array([r.model_prices for r in res])
    market = np.array([r.market_prices for r in res])
    assert rel_err(model, g["model_prices"]).max() <= 1e-10
    assert rel_err(market, g["market_prices"]).max() <= 1e-10
    assert np.array_equal(np.array([r.spot for r in res]), g["spots"])
    assert np.array_equal(np.array([[r.parameters[n] for n in c] for r, c in zip(res, [list(res[0].parameters)] * 20)]),
                          g["params"])
    assert [r.date for r in res] == list(g["dates"])
    assert np.abs(np.array([r.final_loss for r in res]) - g["losses"]).max() <= 1e-12
    assert res[0].calibration_time is None and res[0].iterations is None
    assert res[0].message == str(g["messages"][0])
    o = res[3].market_options[7]
    assert o["option_type"] == "call" and o["maturity"] == 0.5 and o["strike"] == g["strikes"][3, 7]
    with open(path, "rb") as f:
        again = pickle.load(f)
    assert len(again) == 20 and again[5].spot == res[5].spot
    assert gen.generate_synthetic_calibrations(0, path) == []


def test_calibrate_many(mods, golden):
    """C5 shape: many markets x 3 starts in lock-step on the batched optimiser, one launch per round."""
    import dhj
    from oracle import cos_oracle as O
    _, cal, _ = mods
    rng = np.random.default_rng(11)
    n = 24
    params = rng.uniform(O.GENERATOR_RANGES[:, 0], O.GENERATOR_RANGES[:, 1], size=(n, 13))
    spots = rng.uniform(90, 110, size=n)
    K = np.tile(O.GENERATOR_STRIKES_REL[None, :] * spots[:, None] / 100.0, (1, 3))
    T = np.repeat(O.GENERATOR_MATURITIES, 5)
    ctx = dhj.default_context()
    market = ctx.price_list(params, spots, K, T, np.ones(15), 0.03) * (1 + 0.005 * rng.standard_normal((n, 15)))
    np.random.seed(3)
    t0 = time.perf_counter()
    res = dhj.calibrate_many(spots, 0.03, K, T, np.ones(15), market, maxiter=300, multi_start=3, return_all_starts=True)
    wall = time.perf_counter() - t0
    print(f"calibrate_many: {n} markets x 3 starts in {wall:.3f} s, {res['rounds']} launches, "
          f"median loss {np.median(res['final_loss']):.3e}, max {res['final_loss'].max():.3e}")
    assert res['x'].shape == (n, 13) and res['model_prices'].shape == (n, 15)
    assert (res['final_loss'] * 100 < 1.0).all()                    # the reference suite's criterion, every market
    assert np.array_equal(res['final_loss'], res['all_loss'].min(axis=1))
    # fit quality: the calibrated model reprices its own market within the noise that was added
    assert (np.abs(res['model_prices'] - market) / market).max() < 0.05
    # the same markets one at a time through the drop-in calibrator (scipy's L-BFGS-B): same starting points,
    # comparable optima (the two optimisers differ in rounding only; trajectories are chaotic: DESIGN.md §4)
    np.random.seed(3)
    singles = []
    for i in range(6):
        opts = [{'strike': K[i, j], 'maturity': T[j], 'price': market[i, j], 'option_type': 'call'} for j in range(15)]
        c = cal.DoubleHestonJumpCalibrator(spots[i], 0.03, opts)
        singles.append(c.calibrate(maxiter=300, multi_start=3).final_loss)
    singles = np.array(singles)
    ratio = res['final_loss'][:6] / singles
    print("batched / scipy final-loss ratio:", np.round(ratio, 3))
    assert (ratio < 3.0).all() and (ratio > 1 / 3.0).all()
    # the launch count is that of the slowest state, not the sum over states
    assert res['rounds'] <= 21 * 301


def test_calibrate_many_pipelines_agree(mods):
    """Two host pipelines (two contexts, two threads) give the same bits as one lock-step loop: every optimiser
    state is independent of how the states are grouped into launches."""
    import dhj
    from oracle import cos_oracle as O
    rng = np.random.default_rng(17)
    n = 2400
    params = rng.uniform(O.GENERATOR_RANGES[:, 0], O.GENERATOR_RANGES[:, 1], size=(n, 13))
    spots = rng.uniform(90, 110, size=n)
    K = np.tile(O.GENERATOR_STRIKES_REL[None, :] * spots[:, None] / 100.0, (1, 3))
    T = np.repeat(O.GENERATOR_MATURITIES, 5)
    market = dhj.default_context().price_list(params, spots, K, T, np.ones(15), 0.03) * \
        (1 + 0.01 * rng.standard_normal((n, 15)))
    out = []
    for pipes in (1, 2):
        np.random.seed(5)
        t0 = time.perf_counter()
        out.append(dhj.calibrate_many(spots, 0.03, K, T, np.ones(15), market, maxiter=40, multi_start=3,
                                      pipelines=pipes))
        print(f"calibrate_many, {n} markets, maxiter 40, pipelines={pipes}: {time.perf_counter() - t0:.3f} s")
    for key in ('x', 'final_loss', 'iterations', 'status', 'best_start', 'model_prices'):
        assert np.array_equal(out[0][key], out[1][key], equal_nan=True), key


def test_puts_parity_and_jump_limit(mods):
    """The checks the reference's docs promise but its suite does not run (docs/METHODOLOGY.md:150-156):
    put-call parity across the grid, the lambda -> 0 limit, and price bounds."""
    dh, _, _ = mods
    import dhj
    ctx = dhj.default_context()
    p = np.array([TS[k] for k in ("v01", "kappa1", "theta1", "sigma1", "rho1", "v02", "kappa2", "theta2", "sigma2",
                                  "rho2", "lambda_j", "mu_j", "sigma_j")])
    K = np.tile([80.0, 90.0, 100.0, 110.0, 120.0], 3); T = np.repeat([0.25, 1.0, 2.0], 5)
    r = 0.05
    call = ctx.price_list(p, 100.0, K, T, np.ones(15), r)[0]
    put = ctx.price_list(p, 100.0, K, T, np.zeros(15), r)[0]
    parity = call - put - (100.0 - K * np.exp(-r * T))
    assert np.abs(parity).max() < 1e-3                       # the demo's tolerance is 0.01 (double_heston.py:299)
    assert (call > np.maximum(100.0 - K * np.exp(-r * T), 0) - 1e-9).all() and (call < 100.0).all()
    # lambda -> 0: continuous approach to the no-jump price (13.480108655594977 at lambda = 0)
    no_jump = p.copy(); no_jump[10:] = 0.0
    base = ctx.price_list(no_jump, 100.0, [100.0], [1.0], [1], r)[0, 0]
    assert rel_err(base, 13.480108655594977) <= 1e-10
    prev = None
    for lam in (1e-2, 1e-4, 1e-8):
        q = p.copy(); q[10] = lam
        v = ctx.price_list(q, 100.0, [100.0], [1.0], [1], r)[0, 0]
        assert abs(v - base) < 2 * lam * 5.0
        assert prev is None or abs(v - base) <= abs(prev - base)
        prev = v


def test_warm_start_hook_and_flat_arrays(mods, golden, tmp_path):
    _, cal, gen = mods
    gi = golden("initial_guess.npz")
    c = c1_calibrator(cal, gi)
    x_true = c.inverse_transform_params(dict(zip(c.param_names, [0.04, 2.0, 0.04, 0.3, -0.5, 0.04, 1.5, 0.04, 0.2, -0.3,
                                                                 0.1, 0.0, 0.1])))
    res = c.calibrate(maxiter=50, multi_start=1, x0=x_true[None, :] + 0.01)
    assert res.final_loss < 1e-5
    np.random.seed(42)
    path = str(tmp_path / "flat.npz")
    data = gen.generate_synthetic_arrays(20, save_path=path)
    g = golden("generator_seed42.npz")
    assert rel_err(data["model_prices"], g["model_prices"]).max() <= 1e-10
    again = np.load(path)
    assert np.array_equal(again["market_prices"], data["market_prices"]) and again["params"].shape == (20, 13)
    # directory-of-.npy sink + lazy CalibrationResult view against the object form
    np.random.seed(42)
    objs = gen.generate_synthetic_calibrations(20, save_path=str(tmp_path / "objs.pkl"))
    np.random.seed(42)
    gen.generate_synthetic_arrays(20, save_path=str(tmp_path / "flat_dir"))
    ds = gen.SyntheticCalibrationSet.load(tmp_path / "flat_dir")
    assert len(ds) == 20 and isinstance(ds.data["params"], np.memmap)
    for i in (0, 7, 19, -1):
        a_, b_ = ds[i], objs[i]
        assert a_.date == b_.date and a_.spot == b_.spot and a_.parameters == b_.parameters
        assert np.array_equal(a_.market_prices, b_.market_prices) and a_.final_loss == b_.final_loss
        assert a_.market_options == b_.market_options and a_.message == b_.message
    assert ds[5:9][1].date == objs[6].date and len(ds[5:9]) == 4
    assert gen.SyntheticCalibrationSet.load(path)[3].spot == objs[3].spot


def test_reference_suite_section_4_verbatim(mods, golden):
    """tests/test_suite.py:305-344 as written there: scipy drives `compute_loss` itself (jac=None, its own forward
    differences: 14 one-launch loss evaluations per step), maxiter=200, ftol=1e-9; pass if fun*100 < 1."""
    from scipy.optimize import minimize
    _, cal, _ = mods
    gi = golden("initial_guess.npz")
    c = c1_calibrator(cal, gi)
    x0 = c.get_initial_guess()
    result = minimize(fun=c.compute_loss, x0=x0, method='L-BFGS-B', options={'maxiter': 200, 'ftol': 1e-9})
    assert result.fun * 100 < 1.0
    # the reference itself stops at iteration 0 here (ABNORMAL, 294 evaluations, f = 9.76104e-05: BASELINE.md §2)
    ref = golden("calib_trajectory.npz")
    print("suite 4.1 on the GPU: fun %.6e nit %d nfev %d | reference: fun %.6e nit %d nfev %d"
          % (result.fun, result.nit, result.nfev, float(ref["s0_fun"]), int(ref["s0_nit"]), len(ref["s0_fs"])))
    assert abs(result.fun - float(ref["s0_fun"])) <= 1e-6 * float(ref["s0_fun"]) or result.fun < float(ref["s0_fun"])
    assert c.n_calls == result.nfev
    calibrated = c.transform_params(result.x)
    for name, (lo, hi) in {'v1_0': (0.001, 0.5), 'kappa1': (0.1, 10.0), 'theta1': (0.001, 0.5), 'sigma1': (0.01, 2.0),
                           'rho1': (-1.0, 1.0), 'lambda_j': (0.0, 5.0), 'sigma_j': (0.001, 1.0)}.items():
        assert lo <= calibrated[name] <= hi
